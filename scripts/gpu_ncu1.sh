set -x
timeout 600 python -m pytest tests -m gpu -x -q --timeout=150 -k "not full_size" 2>&1 | tail -15
export SNFB_BENCH_SCALE=0.25
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_cluster_warp|k_scan$|k_align|k_vote" -c 6 -o gpurun_out/r2_ncu1 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_ncu1.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/r2_ncu1.log
