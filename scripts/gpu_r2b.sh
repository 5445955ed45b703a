#!/bin/bash
# round-2 session-2 call 1: gpu tests, per-role wait counters (shard-sized and full), A/B on the shard, ncu of the shard-sized scan
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2b_tests.log
( timeout 300 python scripts/dbg_tc_roles.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/r2b_roles_shard.json 2> gpurun_out/r2b_roles_shard.err
( timeout 300 python scripts/dbg_tc_roles.py cfg4_10m_x128_4096b 8 ) > gpurun_out/r2b_roles_full.json 2> gpurun_out/r2b_roles_full.err
( TC_VARIANTS="NLSH_TC_SLOTS=6;NLSH_TC_GRID=140;NLSH_TC_QGLOBAL=0" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/r2b_variants_shard.jsonl 2> gpurun_out/r2b_variants_shard.err
( TC_VARIANTS="NLSH_TC_SLOTS=6" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 ) > gpurun_out/r2b_variants_full.jsonl 2> gpurun_out/r2b_variants_full.err
CMD="python bench.py --workload cfg4_10m_x128_4096b --rows 1250000 --probes 8 --steps 2 --warmup 3 --batches-per-step 2 --no-cpu-baseline"
$CMD > gpurun_out/r2b_plain_shard.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "nlsh_steps/" -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/scan_tc_r2b_shard $CMD > gpurun_out/r2b_ncu_shard.log 2>&1
tail -n 3 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_roles_shard.json gpurun_out/r2b_roles_full.json gpurun_out/r2b_variants_shard.jsonl gpurun_out/r2b_variants_full.jsonl | cut -c1-900; tail -n 2 gpurun_out/r2b_ncu_shard.log
