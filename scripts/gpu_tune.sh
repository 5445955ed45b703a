#!/bin/bash
# A/B of scan-kernel staging variants on one workload (kernel ms from the library's event ring).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg4_10m_x128_4096b}
: > gpurun_out/tune.log
for t in ${TUNES:-"1,3,3" "2,3,2" "1,4,3" "1,2,3" "1,4,2"}; do
  echo "== NLSH_SCAN_TUNE=$t" >> gpurun_out/tune.log
  ( NLSH_SCAN_TUNE=$t timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline ) 2>&1 | \
    python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); r = d['roofline']
        print('QPS %.0f ms/step %.3f kernel_ms %.3f fp32 %.3f hbm_case %.0f GB/s recall %.4f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['fp32_lane_ops_frac'], r['hbm_bound_case']['achieved'], d['config']['recall_at_k']))
    elif 'rror' in line: print(line.rstrip())
" >> gpurun_out/tune.log
done
cat gpurun_out/tune.log
