#!/bin/bash
# Round-2 captures on the default bench command (config 5, one B200); run under gpurun, each only after the plain command exited 0.
#   r2_cfg5_prior:   ncu --set full of the prior-phase kernels (gene-major build, moments, fit version 2)
#   r2_cfg5_sample:  ncu --set full of one column chunk of the 3-D sampler (classification, class kernels, main kernel)
#   r2_cfg5_launches.csv: gpu__time_duration.sum of every launch of one step
set -x
WHAT=${1:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = prior ]; then
ncu --set full --import-source on --clock-control none -k regex:"k_gm_scatter|k_gm_count|k_row_moments|k_fit2|k_posterior_w" -c 8 -o gpurun_out/r2_cfg5_prior -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2a.log 2>&1
fi
if [ "$WHAT" = all ] || [ "$WHAT" = sample ]; then
ncu --set full --import-source on --clock-control none -k regex:"k_post_sample_main|k_nz_|k_post_sample_heavy" -c 10 -o gpurun_out/r2_cfg5_sample -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2b.log 2>&1
fi
if [ "$WHAT" = all ] || [ "$WHAT" = launches ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 21000 --csv --log-file gpurun_out/r2_cfg5_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2c.log 2>&1
fi
