#!/bin/bash
# ncu --set full of the index-build kernels (one nlsh_build_csr call, 10M x 128, 4096 buckets).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
python scripts/build_profile.py > gpurun_out/build_profile_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
  -k "regex:unit_hist|unit_scatter|gather_rows|scan_reduce|scan_sums|scan_apply|extract_offsets" -s 7 -c 7 \
  -f -o gpurun_out/build_$TAG python scripts/build_profile.py > gpurun_out/ncu_build_$TAG.log 2>&1
tail -n 2 gpurun_out/build_profile_$TAG.log gpurun_out/ncu_build_$TAG.log
