set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final3_c2.json 2> gpurun_out/r2_final3_c2.err; echo rc=$?
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_final3_c3.json 2> gpurun_out/r2_final3_c3.err; echo rc=$?
timeout 600 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2_final3_c6.json 2> gpurun_out/r2_final3_c6.err; echo rc=$?
python - <<'PY'
import json
for c in (2,3):
    d=json.load(open(f"gpurun_out/r2_final3_c{c}.json"))
    print(c,"value",d["value"],"ms",d["ms_per_step"],"dev",d.get("device_ms_per_step"),"parity",d.get("parity_full_size"),"hash",d.get("callset_sha256")[:12])
    print({k:round(v,3) for k,v in d["stage_ms"].items()})
d=json.load(open("gpurun_out/r2_final3_c6.json")); print(d["value"], d["stage_ms"], d["e2e"], d.get("cpu_baseline"))
PY
