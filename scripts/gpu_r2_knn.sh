#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2k}
( timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_tc.py tests/test_gpu_hash.py -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
( timeout 300 python scripts/knn_time.py ) > gpurun_out/${T}_knn_time.txt 2>&1
( timeout 300 python scripts/bench_hasher.py ) > gpurun_out/${T}_bench_hasher.json 2>&1
tail -n 3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_knn_time.txt; cut -c1-700 gpurun_out/${T}_bench_hasher.json
