#!/bin/bash
# One GPU-box pass: smoke, the gpu test files (one process each so a CUDA fault in one file
# does not poison the rest), then short bench runs.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
for f in tests/test_gpu_hash.py tests/test_gpu_tc.py tests/test_gpu_build.py tests/test_gpu_query.py tests/test_gpu_knn.py tests/test_gpu_api.py; do
  name=$(basename $f .py)
  ( timeout 900 python -m pytest $f -m gpu -q -x --timeout=600 -s ) > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
done
for w in ${BENCH_WORKLOADS:-cfg2_1m_x128_256b cfg4_10m_x128_4096b}; do
  ( timeout 900 python bench.py --workload $w --steps 10 --warmup 3 ) > gpurun_out/bench_$w.log 2>&1
  echo "exit $?" >> gpurun_out/bench_$w.log
done
tail -n 5 gpurun_out/*.log
