set -x
# state of the tree: GPU tests, then one bench line per config (CUDA-event timings, outside any profiler)
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_c2.json 2> gpurun_out/r2_final_c2.err; echo rc=$?
for c in 1 5 4; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_final_c$c.json 2> gpurun_out/r2_final_c$c.err; echo rc=$?; done
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_final_c3.json 2> gpurun_out/r2_final_c3.err; echo rc=$?
tail -3 gpurun_out/*.err
