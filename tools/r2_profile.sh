set -x
ncu --set full --import-source on --clock-control none -k regex:"k_gm_scatter|k_row_moments|k_fit2|k_posterior_w" -c 6 -o gpurun_out/r2_cfg5_prior -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2a.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_post_sample_main|k_nz_|k_post_sample_heavy" -c 10 -o gpurun_out/r2_cfg5_sample -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 21000 --csv --log-file gpurun_out/r2_cfg5_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2c.log 2>&1
