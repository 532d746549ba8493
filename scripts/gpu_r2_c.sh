set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -5
for c in 1 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_c$c.json 2> gpurun_out/r2_c$c.err; echo rc=$?; done
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_c3.json 2> gpurun_out/r2_c3.err; echo rc=$?
