#!/bin/bash
# ncu --set full capture of the tensor-core filtered scan kernel on the bench workload.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
TAG=${2:-r1f}
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s ${SKIP:-16} -c 1 -f -o gpurun_out/scan_tc_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/plain_$TAG.log gpurun_out/ncu_full_$TAG.log
