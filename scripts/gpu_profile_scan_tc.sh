#!/bin/bash
# ncu passes for the bench with the tensor-core filtered scan (1 GPU): launch list of the eager
# query steps (NVTX range "nlsh_steps" in bench.py) + one full capture of scan_tc_kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
TAG=${2:-r1g}
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "nlsh_steps/" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s ${SKIP:-16} -c 1 -f -o gpurun_out/scan_tc_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 2 gpurun_out/plain_$TAG.log | cut -c1-400; tail -n 3 gpurun_out/ncu_launches_$TAG.log gpurun_out/ncu_full_$TAG.log | cut -c1-300
