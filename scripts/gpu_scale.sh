#!/bin/bash
# bench at N ranks only (scaling smoke for N > 2): gpurun --gpus N -- bash scripts/gpu_scale.sh N tag
N=${1:-4}; TAG=${2:-scale}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpus.txt
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/${TAG}_bench_n$N.json | cut -c1-600
echo "== reference arm at N=$N (rank 0 only runs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_reference_n$N.json | cut -c1-300
