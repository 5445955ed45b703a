#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2g}
L=$PWD/neural-locality-sensitive-hashing_b200/lib
( timeout 600 python -m pytest tests/test_gpu_query.py tests/test_gpu_api.py -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
V="NLSH_TC_SETS=16;NLSH_TC_SETS=4;NLSH_TC_SETS=2"
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_variants_shard.jsonl 2> gpurun_out/${T}_variants_shard.err
( NLSH_B200_LIB=$L/libnlsh_b200_sb4.so TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/${T}_variants_shard_sb4.jsonl 2> gpurun_out/${T}_variants_shard_sb4.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 ) > gpurun_out/${T}_variants_full.jsonl 2> gpurun_out/${T}_variants_full.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg3_1.2m_x100_1024b_angular 2 ) > gpurun_out/${T}_variants_cfg3.jsonl 2> gpurun_out/${T}_variants_cfg3.err
( NLSH_B200_LIB=$L/libnlsh_b200_sb4.so TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg3_1.2m_x100_1024b_angular 2 ) > gpurun_out/${T}_variants_cfg3_sb4.jsonl 2> gpurun_out/${T}_variants_cfg3_sb4.err
( TC_VARIANTS="$V" timeout 300 python scripts/dbg_tc_variants.py cfg2_1m_x128_256b 4 ) > gpurun_out/${T}_variants_cfg2.jsonl 2> gpurun_out/${T}_variants_cfg2.err
( TC_VARIANTS="NLSH_TC_SETS=4;NLSH_TC_SETS=2" timeout 300 python scripts/dbg_tc_variants.py cfg5_1m_x960_512b_k100 128 ) > gpurun_out/${T}_variants_cfg5.jsonl 2> gpurun_out/${T}_variants_cfg5.err
tail -n 3 gpurun_out/${T}_tests.log; for f in shard shard_sb4 full cfg3 cfg3_sb4 cfg2 cfg5; do echo "== $f"; cut -c1-250 gpurun_out/${T}_variants_$f.jsonl; tail -n 2 gpurun_out/${T}_variants_$f.err; done
