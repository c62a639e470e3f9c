#!/bin/bash
# multi-GPU visit (gpurun --gpus N): NCCL grid tests, config C4 measured at 1 and N ranks, the bench at N ranks (and N=1)
N=${1:-2}; TAG=${2:-multi}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpus.txt
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -5 | tee $O/${TAG}_pytest.txt
for n in 1 $N; do
echo "== C4 at $n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
   scripts/measure_c4.py 2>&1 | tail -1 | tee $O/${TAG}_c4_n$n.json | cut -c1-600
done
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/${TAG}_bench_n1.json | cut -c1-400
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/${TAG}_bench_n$N.json | cut -c1-400
