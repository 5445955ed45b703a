#!/bin/bash
# final single-GPU pass of the round: gpu tests, smoke, the bench line of every workload, launch list + full ncu capture of cfg4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2z}
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1; echo "exit $?" >> gpurun_out/${T}_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1; echo "exit $?" >> gpurun_out/${T}_smoke.log
for w in cfg4_10m_x128_4096b cfg1_100k_x128_16b cfg2_1m_x128_256b cfg3_1.2m_x100_1024b_angular cfg5_1m_x960_512b_k100 cfg4_10m_x128_4096b_skewed; do
  ( timeout 600 python bench.py --workload $w ) > gpurun_out/${T}_bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/${T}_bench_$w.log
done
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_reference_arm.log 2>&1
FULL=1 bash scripts/gpu_profile_scan_tc.sh cfg4_10m_x128_4096b $T > gpurun_out/${T}_profile.log 2>&1
tail -n 2 gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log; for w in cfg4_10m_x128_4096b cfg1_100k_x128_16b cfg2_1m_x128_256b cfg3_1.2m_x100_1024b_angular cfg5_1m_x960_512b_k100 cfg4_10m_x128_4096b_skewed; do grep '^{' gpurun_out/${T}_bench_$w.log | cut -c1-200; done; tail -n 1 gpurun_out/${T}_bench_reference_arm.log | cut -c1-300; tail -n 4 gpurun_out/${T}_profile.log | cut -c1-300
