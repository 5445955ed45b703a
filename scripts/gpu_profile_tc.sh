#!/bin/bash
# ncu --set full captures of the two tensor-core kernels (hasher layer, brute-force kNN).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
python scripts/knn_profile.py 10 > gpurun_out/knn_profile_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_knn_kernel -s 2 -c 1 -f -o gpurun_out/tc_knn_$TAG python scripts/knn_profile.py 10 > gpurun_out/ncu_tc_knn_$TAG.log 2>&1
python scripts/bench_hasher.py 2000000 > gpurun_out/hasher_profile_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 6 -c 3 -f -o gpurun_out/tc_linear_$TAG python scripts/bench_hasher.py 2000000 > gpurun_out/ncu_tc_linear_$TAG.log 2>&1
tail -n 3 gpurun_out/knn_profile_$TAG.log gpurun_out/ncu_tc_knn_$TAG.log gpurun_out/hasher_profile_$TAG.log gpurun_out/ncu_tc_linear_$TAG.log
