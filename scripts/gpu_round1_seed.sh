#!/bin/bash
# Bucket-ordered seed kernel (NLSH_SEED_ORDER=bucket): every test that reaches the tensor-core scan
# runs with it, then a run-time A/B on the full database and on one 8-GPU shard's worth of rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( NLSH_SEED_ORDER=bucket timeout 300 python -m pytest tests/test_gpu_query.py tests/test_gpu_api.py -m gpu -q -x ) > gpurun_out/test_seed_order.log 2>&1
echo "exit $?" >> gpurun_out/test_seed_order.log
( TC_VARIANTS="NLSH_SEED_ORDER=bucket;NLSH_SEED_ORDER=bucket,NLSH_SCAN_SEED=384" \
  timeout 200 python scripts/dbg_tc_variants.py ) > gpurun_out/variants_order_full.log 2>&1
echo "exit $?" >> gpurun_out/variants_order_full.log
( TC_VARIANTS="NLSH_SEED_ORDER=bucket;NLSH_SEED_ORDER=bucket,NLSH_SCAN_SEED=192;NLSH_SEED_ORDER=bucket,NLSH_SCAN_SEED=256" \
  timeout 200 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/variants_order_shard.log 2>&1
echo "exit $?" >> gpurun_out/variants_order_shard.log
tail -n 3 gpurun_out/test_seed_order.log; cut -c1-230 gpurun_out/variants_order_full.log gpurun_out/variants_order_shard.log
