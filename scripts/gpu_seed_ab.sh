#!/bin/bash
# Pipelined seed kernel (three row steps in flight; NLSH_SEED_PIPE=0 selects the old one): every test
# that reaches the tensor-core scan, then a run-time A/B on the full database and on one 8-GPU
# shard's worth of rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_query.py tests/test_gpu_api.py -m gpu -q -x ) > gpurun_out/test_seed_pipe.log 2>&1
echo "exit $?" >> gpurun_out/test_seed_pipe.log
( TC_VARIANTS="NLSH_SEED_PIPE=0;NLSH_SCAN_SEED=384" \
  timeout 200 python scripts/dbg_tc_variants.py ) > gpurun_out/variants_pipe_full.log 2>&1
echo "exit $?" >> gpurun_out/variants_pipe_full.log
( TC_VARIANTS="NLSH_SEED_PIPE=0;NLSH_SCAN_SEED=192;NLSH_SCAN_SEED=256" \
  timeout 200 python scripts/dbg_tc_variants.py cfg4_10m_x128_4096b 8 1250000 ) > gpurun_out/variants_pipe_shard.log 2>&1
echo "exit $?" >> gpurun_out/variants_pipe_shard.log
tail -n 3 gpurun_out/test_seed_pipe.log; cut -c1-230 gpurun_out/variants_pipe_full.log gpurun_out/variants_pipe_shard.log
