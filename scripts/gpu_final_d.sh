set -x
timeout 200 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2_final4_c6.json 2> gpurun_out/r2_final4_c6.err; echo rc=$?
tail -2 gpurun_out/r2_final4_c6.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_final4_c6.json")); print(d["value"], d["stage_ms"], d["e2e"], d.get("cpu_baseline"), d["roofline"]["frac"])
PY
