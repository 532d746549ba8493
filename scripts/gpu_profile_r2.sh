set -x
# launch list of the bench command (serialised, cold-cache per-launch times: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_launches_full.log 2>&1; echo rc=$?
# one full capture of the kernels that carry the step
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_scan$|k_align|k_vote|k_cluster_warp|k_rec_index|k_coverage" -c 8 -o gpurun_out/r2_ncu_full python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/ | tail -5
