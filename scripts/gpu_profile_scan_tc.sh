#!/bin/bash
# ncu passes for the bench with the tensor-core filtered scan (1 GPU): launch list of the eager
# query steps (NVTX range "nlsh_steps" in bench.py) + one full capture of scan_tc_kernel.
# Each ncu pass runs only after the same command has exited 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
TAG=${2:-r2a}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --batches-per-step 2 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "nlsh_steps/" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
if [ "${FULL:-1}" = "1" ]; then
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "nlsh_steps/" -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/scan_tc_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
fi
tail -n 2 gpurun_out/plain_$TAG.log | cut -c1-400; tail -n 3 gpurun_out/ncu_launches_$TAG.log gpurun_out/ncu_full_$TAG.log | cut -c1-300
