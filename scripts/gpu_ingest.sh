set -x
timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_parity.py::test_calltask_from_bam_region -x -q --timeout=300 2>&1 | tail -25
timeout 600 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2_ingest_c6.json 2> gpurun_out/r2_ingest_c6.err; echo rc=$?
tail -5 gpurun_out/r2_ingest_c6.err; cat gpurun_out/r2_ingest_c6.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_inflate" -c 1 -o gpurun_out/r2_ncu_inflate python bench.py --config 6 --scale 0.5 --ingest-tiles 4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_inflate.log 2>&1; echo rc=$?
