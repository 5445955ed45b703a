#!/bin/bash
# 8-GPU pass: the bench line, then the lanes / reserve sweep of the serving loop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/r2i_bench_n$N.log
( LANE_VARIANTS="${LANES:-4:0;4:16;6:0;6:16;8:16}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/dbg_lanes_multi.py ) > gpurun_out/r2i_lanes_n$N.log 2>&1
echo "exit $?" >> gpurun_out/r2i_lanes_n$N.log
grep '^{' gpurun_out/r2i_bench_n$N.log | cut -c1-2500; tail -n 2 gpurun_out/r2i_bench_n$N.log | cut -c1-300; grep '^{' gpurun_out/r2i_lanes_n$N.log; tail -n 2 gpurun_out/r2i_lanes_n$N.log | cut -c1-300
