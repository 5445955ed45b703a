#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2s}
V="NLSH_SCAN_SEED_DIV=24;NLSH_SCAN_SEED_DIV=32;NLSH_SCAN_SEED_DIV=12;NLSH_SCAN_SEED_DIV=48"
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 ) > gpurun_out/${T}_variants_full.jsonl 2> gpurun_out/${T}_variants_full.err
V="NLSH_SCAN_SEED=64;NLSH_SCAN_SEED=96;NLSH_SCAN_SEED=160;NLSH_SCAN_SEED=192"
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_variants_shard.jsonl 2> gpurun_out/${T}_variants_shard.err
for f in full shard; do echo "== $f"; cut -c1-260 gpurun_out/${T}_variants_$f.jsonl; tail -n 2 gpurun_out/${T}_variants_$f.err; done
