set -x
timeout 300 python -m pytest tests -m gpu -x -q --timeout=200 -k "parity and not full" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_b7_n1.json 2> gpurun_out/r2_b7_n1.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/r2_b7_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['parity_vs_n1']['identical'], {k:round(v,3) for k,v in d['stage_ms'].items()})"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_align|k_vote|k_prep" --launch-skip 3 -c 3 -o gpurun_out/r2_ncu_consensus python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2_ncu_consensus.log 2>&1; echo rc=$?
