#!/bin/bash
# final checks: the whole gpu suite on the final library, then compute-sanitizer memcheck over the scan tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2y_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2y_tests.log
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_query.py -m gpu -x -q -k "filter or tc or shard or overflow or wide" ) > gpurun_out/r2y_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r2y_memcheck.log
tail -n 3 gpurun_out/r2y_tests.log; grep -c "Invalid\|out of bounds" gpurun_out/r2y_memcheck.log; tail -n 6 gpurun_out/r2y_memcheck.log | cut -c1-200
