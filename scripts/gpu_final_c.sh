set -x
timeout 300 python -m pytest tests -m gpu -x -q --timeout=120 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2_final4_c2.json 2> gpurun_out/r2_final4_c2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_final4_c2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"dev",d.get("device_ms_per_step"),"hash",d.get("callset_sha256")[:12])
print({k:round(v,3) for k,v in d["stage_ms"].items()})
PY
