#!/bin/bash
# Launch list of the eager query steps (NVTX range "nlsh_steps" in bench.py) with WARM caches
# (--cache-control none): the small kernels of a step read what the previous kernel just wrote.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
TAG=${2:-r2w}
ROWS=${3:-0}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --batches-per-step 2 --no-cpu-baseline --rows $ROWS"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --nvtx --nvtx-include "nlsh_steps/" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -n 1 gpurun_out/plain_$TAG.log | cut -c1-300
