#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2j_bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/r2j_bench_n$N.log
grep '^{' gpurun_out/r2j_bench_n$N.log | head -1 | cut -c1-2600; tail -n 3 gpurun_out/r2j_bench_n$N.log | cut -c1-300
