#!/bin/bash
# round 2, call 26: source-level (SASS + CUDA line) instruction profile of the two search kernels
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:self_knn_pca -c 1 -f -o /tmp/cov python tools/probe_knn_once.py > gpurun_out/r2_src_cov.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cross_knn -s 1 -c 1 -f -o /tmp/knn python tools/probe_knn_once.py > gpurun_out/r2_src_knn.log 2>&1
for k in cov knn; do
  ncu -i /tmp/$k.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_src_${k}_sass.csv 2>/dev/null
  ncu -i /tmp/$k.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2_src_${k}_cuda.csv 2>/dev/null
done
ls -la gpurun_out/r2_src_*
