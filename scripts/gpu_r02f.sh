#!/bin/bash
# r02f visit (8 GPUs): the multi-GPU tests at full width, in-process e2e over 8 GPUs, bench at N=8 and N=1 on one box
TAG=r02f
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest multi-GPU"; timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multidev.py -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest_multi.txt; tail -15 $O/${TAG}_pytest_multi.txt | cut -c1-300
echo "== e2e sweep in-process, 8 GPUs"; timeout 500 python scripts/e2e_scaling.py --devices 8 --modes spin --threads 16,24,32,48,64 --seconds 1.5 > $O/${TAG}_e2e_inproc8.txt 2>&1; cat $O/${TAG}_e2e_inproc8.txt
echo "== bench N=8"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; tail -c 1200 $O/${TAG}_bench_n8.json; tail -5 $O/${TAG}_bench_n8.err
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 300 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-300
ls -la $O | tail -8
