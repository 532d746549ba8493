set -x
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo rc=$?
grep -v "^\[bench\]" gpurun_out/r2_scale_n$N.err | tail -5
