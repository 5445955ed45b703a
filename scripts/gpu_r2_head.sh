#!/bin/bash
# fused output layer in the epilogue of the layer before it: tests, hasher timing with / without, a query batch with / without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2u
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
( timeout 200 python scripts/bench_hasher.py ) > gpurun_out/${T}_bench_hasher_fused.json 2>&1
( NLSH_MLP_FUSE_HEAD=0 timeout 200 python scripts/bench_hasher.py ) > gpurun_out/${T}_bench_hasher_unfused.json 2>&1
( timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/${T}_bench_n1.log 2>&1
tail -n 3 gpurun_out/${T}_tests.log; cut -c1-420 gpurun_out/${T}_bench_hasher_fused.json; echo; cut -c1-420 gpurun_out/${T}_bench_hasher_unfused.json; echo; grep '^{' gpurun_out/${T}_bench_n1.log | cut -c1-200; grep -o '"breakdown_ms_per_batch": {[^}]*' gpurun_out/${T}_bench_n1.log | cut -c1-300
