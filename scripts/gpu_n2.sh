set -x
timeout 300 python -m pytest tests -m gpu -x -q --timeout=150 -k "parity and not golden" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_b4_n1.json 2> gpurun_out/r2_b4_n1.err; echo rc=$?
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b4_n2.json 2> gpurun_out/r2_b4_n2.err; echo rc=$?
tail -5 gpurun_out/r2_b4_n2.err
