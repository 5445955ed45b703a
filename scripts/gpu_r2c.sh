#!/bin/bash
# planner / streamer split + hoisted scorer loads: tests, per-role counters, A/B on every config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2c}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
( timeout 300 python scripts/dbg_tc_roles.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_roles_shard.json 2> gpurun_out/${T}_roles_shard.err
( timeout 300 python scripts/dbg_tc_roles.py cfg4_10m_x128_4096b 8 ) > gpurun_out/${T}_roles_full.json 2> gpurun_out/${T}_roles_full.err
V=${VARS:-"NLSH_TC_SLOTS=4;NLSH_TC_SLOTS=5;NLSH_TC_SLOTS=6;NLSH_TC_SLOTS=7;NLSH_TC_V8=0;NLSH_TC_SLOTS=6,NLSH_TC_V8=0"}
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_variants_shard.jsonl 2> gpurun_out/${T}_variants_shard.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 ) > gpurun_out/${T}_variants_full.jsonl 2> gpurun_out/${T}_variants_full.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg2_1m_x128_256b 4 ) > gpurun_out/${T}_variants_cfg2.jsonl 2> gpurun_out/${T}_variants_cfg2.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg3_1.2m_x100_1024b_angular 2 ) > gpurun_out/${T}_variants_cfg3.jsonl 2> gpurun_out/${T}_variants_cfg3.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg5_1m_x960_512b_k100 128 ) > gpurun_out/${T}_variants_cfg5.jsonl 2> gpurun_out/${T}_variants_cfg5.err
tail -n 3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_roles_shard.json gpurun_out/${T}_roles_full.json | cut -c1-1200; for f in shard full cfg2 cfg3 cfg5; do echo "== $f"; cut -c1-330 gpurun_out/${T}_variants_$f.jsonl; tail -n 2 gpurun_out/${T}_variants_$f.err; done
