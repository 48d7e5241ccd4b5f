set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
tail -c 600 gpurun_out/r2_base_bench.json
python tools/trace_batch.py 16 8 > gpurun_out/r2_base_trace.log 2>&1
tail -20 gpurun_out/r2_base_trace.log
timeout 300 ncu --set full --clock-control none -k regex:lm_kernel -s 8 -c 2 -f -o gpurun_out/r2_base_lm_batch python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pairs 8 > gpurun_out/r2_base_ncu.log 2>&1
ncu -i gpurun_out/r2_base_lm_batch.ncu-rep --page raw --csv > gpurun_out/r2_base_lm_batch_raw.csv 2>/dev/null
tail -3 gpurun_out/r2_base_ncu.log
