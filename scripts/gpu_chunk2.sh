set -x
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_chunk2_c2.json 2> gpurun_out/r2_chunk2_c2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_chunk2_c2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"dev",d.get("device_ms_per_step"),"parity",d.get("parity_full_size"),"hash",d.get("callset_sha256"))
print({k:round(v,3) for k,v in d["stage_ms"].items()})
print(d["rooflines"][0])
PY
for L in 32 16 8; do SNFB_INFLATE_LANES=$L timeout 300 python bench.py --config 6 --scale 0.5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_ingest_L$L.json 2> gpurun_out/r2_ingest_L$L.err; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r2_ingest_L$L.json')); print('lanes $L', d['value'], d['stage_ms'])"; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_chunk_sum|k_chunk_rare|k_rec_base|k_rec_fin" -c 4 -o gpurun_out/r2_ncu_chunk2 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_ncu_chunk2.log 2>&1; echo rc=$?
