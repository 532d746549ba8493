set -x
# launch list of the bench command of the final tree (serialised, cold-cache per-launch times: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_launches_full.log 2>&1; echo rc=$?
# the DEFLATE kernel (16 lanes per BGZF block), one launch
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_inflate" -c 1 -o gpurun_out/r2_ncu_inflate16 python bench.py --config 6 --scale 0.5 --ingest-tiles 4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_inflate16.log 2>&1; echo rc=$?
for c in 1 5; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_final2_c$c.json 2> gpurun_out/r2_final2_c$c.err; echo rc=$?; done
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_final2_c3.json 2> gpurun_out/r2_final2_c3.err; echo rc=$?
