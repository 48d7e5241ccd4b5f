# round 2, call 5: why is the Hilbert build slower on the lone kNN kernels?  ncu of both builds + new GPU tests
set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in default morton; do
  lib=$L/libsicp_b200.so; [ $v = morton ] && lib=$L/libsicp_b200_morton.so
  SICP_LIB=$lib timeout 300 ncu --set full --import-source on --clock-control none -k regex:"cross_knn|self_knn" -s 1 -c 2 -f -o gpurun_out/r2_c5_knn_$v python tools/probe_knn_once.py > gpurun_out/r2_c5_ncu_$v.log 2>&1
  ncu -i gpurun_out/r2_c5_knn_$v.ncu-rep --page raw --csv > gpurun_out/r2_c5_knn_${v}_raw.csv 2>/dev/null
  python tools/ncu_pick.py gpurun_out/r2_c5_knn_${v}_raw.csv
done
