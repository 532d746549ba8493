set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_chunk_c2.json 2> gpurun_out/r2_chunk_c2.err; echo rc=$?
tail -3 gpurun_out/r2_chunk_c2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_chunk_c2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"dev",d.get("device_ms_per_step"),"parity",d.get("parity_full_size"),"hash",d.get("callset_sha256"),d.get("callset_matches_committed"))
print({k:round(v,3) for k,v in d["stage_ms"].items()})
print(d["rooflines"][0])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chunk_sum|k_chunk_rare|k_rec_base|k_rec_fin|k_cdesc" -c 5 -o gpurun_out/r2_ncu_chunk python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_ncu_chunk.log 2>&1; echo rc=$?
