# ncu of the LM kernel in the batch configuration (37-CTA grids; pass-by-pass path so that ncu sees individual launches), source-level
set -x
cd "$(dirname "$0")/../.."
SICP_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lm_kernel" -s 12 -c 4 -f -o gpurun_out/r2_lm_batch python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --pairs 8 > gpurun_out/r2_lm_batch_ncu.log 2>&1
ncu -i gpurun_out/r2_lm_batch.ncu-rep --page raw --csv > gpurun_out/r2_lm_batch_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lm_batch.ncu-rep --page source --csv > gpurun_out/r2_lm_batch_src.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/r2_lm_batch_raw.csv gpu__time_duration.sum launch__grid_size sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
