#!/bin/bash
# Bucket-grouped seed kernel: exactness (the tensor-core scan must still equal the SIMT scan bit for
# bit) and run-time A/B on the full database and on one 8-GPU shard's worth of rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( NLSH_SEED_IMPL=bucket timeout 200 python -m pytest tests/test_gpu_query.py -m gpu -q -x -k tensor_core_filter ) > gpurun_out/test_seed_bucket.log 2>&1
echo "exit $?" >> gpurun_out/test_seed_bucket.log
( TC_VARIANTS="NLSH_SEED_IMPL=bucket;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=384;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=512;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=128" \
  timeout 200 python scripts/dbg_tc_variants.py ) > gpurun_out/variants_seed_full.log 2>&1
echo "exit $?" >> gpurun_out/variants_seed_full.log
( TC_VARIANTS="NLSH_SEED_IMPL=bucket;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=192;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=256;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=384;NLSH_SEED_IMPL=bucket,NLSH_SCAN_SEED=512" \
  timeout 200 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/variants_seed_shard.log 2>&1
echo "exit $?" >> gpurun_out/variants_seed_shard.log
( TC_VARIANTS="NLSH_SEED_IMPL=bucket" timeout 200 python scripts/dbg_tc_variants.py cfg2_1m_x128_256b 4 ) > gpurun_out/variants_seed_cfg2.log 2>&1
echo "exit $?" >> gpurun_out/variants_seed_cfg2.log
tail -n 3 gpurun_out/test_seed_bucket.log; cut -c1-250 gpurun_out/variants_seed_full.log gpurun_out/variants_seed_shard.log gpurun_out/variants_seed_cfg2.log
