#!/bin/bash
# One GPU-box pass late in round 1: exactness of the prefetch switches, run-time A/B of the
# tensor-core scan (full database and one 8-GPU shard's worth of rows), ncu of the index build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 2; do
  ( NLSH_TC_PREFETCH=$v timeout 200 python -m pytest tests/test_gpu_query.py -m gpu -q -x -k tensor_core_filter ) > gpurun_out/test_prefetch$v.log 2>&1
  echo "exit $?" >> gpurun_out/test_prefetch$v.log
done
( timeout 200 python scripts/dbg_tc_variants.py ) > gpurun_out/variants_full.log 2>&1
echo "exit $?" >> gpurun_out/variants_full.log
( TC_VARIANTS="NLSH_SCAN_SEED=64;NLSH_SCAN_SEED=256;NLSH_TC_PREFETCH=1;NLSH_TC_PREFETCH=2;NLSH_TC_PREFETCH=1,NLSH_SCAN_SEED=256;NLSH_TC_PREFETCH=1,NLSH_TC_SLOTS=5" \
  timeout 200 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/variants_shard.log 2>&1
echo "exit $?" >> gpurun_out/variants_shard.log
timeout 300 bash scripts/gpu_profile_build.sh r1
tail -n 3 gpurun_out/test_prefetch*.log; cut -c1-330 gpurun_out/variants_full.log gpurun_out/variants_shard.log
