# evidence: GPU tests (incl. the paired path), synccheck, ncu launch list of the bench command, ncu --set full of one lone registration
set -x
cd "$(dirname "$0")/../.."
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 compute-sanitizer --tool synccheck python tools/sanitize_small.py > gpurun_out/r2_san_synccheck.txt 2>&1; tail -12 gpurun_out/r2_san_synccheck.txt
SICP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --pairs 8 > gpurun_out/r2_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches.csv
SICP_GRAPH=0 timeout 600 ncu --set full --clock-control none -f -o /tmp/r2_full python tools/probe_one.py > gpurun_out/r2_full.log 2>&1; tail -2 gpurun_out/r2_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; wc -l gpurun_out/r2_full_raw.csv
