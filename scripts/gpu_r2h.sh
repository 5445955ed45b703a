#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2h}
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
CMD="python bench.py --workload cfg4_10m_x128_4096b --rows 1250000 --probes 8 --steps 2 --warmup 3 --batches-per-step 2 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/${T}_plain_shard.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --nvtx --nvtx-include "nlsh_steps/" -c 400 --csv --log-file gpurun_out/${T}_launches_shard.csv $CMD > gpurun_out/${T}_ncu_launches_shard.log 2>&1
tail -n 3 gpurun_out/${T}_tests.log; tail -n 1 gpurun_out/${T}_plain_shard.log | cut -c1-600; tail -n 3 gpurun_out/${T}_ncu_launches_shard.log
