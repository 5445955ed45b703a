#!/bin/bash
# ncu passes for the bench (1 GPU): launch list of a short run + one full capture of the scan kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
TAG=${2:-r1}
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scan_kernel|plan_|merge_|linear_act|codes_kernel|probes_|unit_hist|unit_scatter|gather_rows|normalize_rows|extract_offsets|recall_hits|fill_int" -s ${LSKIP:-1233} -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s ${SKIP:-16} -c 1 -f -o gpurun_out/scan_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/plain_$TAG.log gpurun_out/ncu_launches_$TAG.log gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
