#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2f}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
( timeout 300 python scripts/dbg_tc_roles.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_roles_shard.json 2> gpurun_out/${T}_roles_shard.err
V=${VARS:-"NLSH_TC_SLOTS=5;NLSH_TC_SLOTS=7;NLSH_TC_SEED=64"}
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_variants_shard.jsonl 2> gpurun_out/${T}_variants_shard.err
( TC_VARIANTS="NLSH_TC_SLOTS=7" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 ) > gpurun_out/${T}_variants_full.jsonl 2> gpurun_out/${T}_variants_full.err
( TC_VARIANTS="NLSH_TC_SLOTS=5" timeout 300 python scripts/dbg_tc_variants.py cfg3_1.2m_x100_1024b_angular 2 ) > gpurun_out/${T}_variants_cfg3.jsonl 2> gpurun_out/${T}_variants_cfg3.err
( TC_VARIANTS="NLSH_TC_SLOTS=5" timeout 300 python scripts/dbg_tc_variants.py cfg5_1m_x960_512b_k100 128 ) > gpurun_out/${T}_variants_cfg5.jsonl 2> gpurun_out/${T}_variants_cfg5.err
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/${T}_bench_n1.log 2>&1
tail -n 3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_roles_shard.json | cut -c1-1300; for f in shard full cfg3 cfg5; do echo "== $f"; cut -c1-300 gpurun_out/${T}_variants_$f.jsonl; tail -n 2 gpurun_out/${T}_variants_$f.err; done; tail -n 1 gpurun_out/${T}_bench_n1.log | cut -c1-1500
