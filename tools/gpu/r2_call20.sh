# round 2: sanitizer, ncu launch list of the bench command, ncu --set full of one lone registration (pass-by-pass path: ncu does not see kernels inside a conditional graph body)
set -x
cd "$(dirname "$0")/../.."
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4 > gpurun_out/r2_san_memcheck.txt; cat gpurun_out/r2_san_memcheck.txt
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_small.py 2>&1 | tail -3 > gpurun_out/r2_san_synccheck.txt; cat gpurun_out/r2_san_synccheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py 2>&1 | tail -3 > gpurun_out/r2_san_racecheck.txt; cat gpurun_out/r2_san_racecheck.txt
SICP_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --pairs 16 > gpurun_out/r2_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/r2_bench_under_ncu.log; wc -l gpurun_out/r2_launches.csv
SICP_GRAPH=0 timeout 900 ncu --set full --import-source on --clock-control none -f -o gpurun_out/r2_full python tools/probe_one.py > gpurun_out/r2_full.log 2>&1; tail -2 gpurun_out/r2_full.log
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; wc -l gpurun_out/r2_full_raw.csv
