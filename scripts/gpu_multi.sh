#!/bin/bash
# N-GPU bench pass (NCCL all-gather + merge); run under `gpurun --gpus N`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
W=${2:-cfg4_10m_x128_4096b}
STEPS=${3:-20}
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $W --steps $STEPS --warmup 5 ) > gpurun_out/bench_${W}_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_${W}_n$N.log
tail -n 4 gpurun_out/bench_${W}_n$N.log
