#!/bin/bash
# r02b visit (2 GPUs): cross-device paths of the in-process pool, NCCL grid tests at C4's size, bench at N=1 and N=2,
# in-process e2e sweep over 2 GPUs, NN-mode timings after the scan/ticket changes.
TAG=r02b
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
nvidia-smi topo -m >> $O/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -x 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== configs (NN after row_scan/ticket changes)"; timeout 600 python scripts/measure_configs.py > $O/${TAG}_configs.txt 2>&1; grep -E "'nn'" $O/${TAG}_configs.txt | cut -c1-230
echo "== e2e sweep in-process, 2 GPUs"; timeout 400 python scripts/e2e_scaling.py --devices 2 --modes spin --threads 8,16,24,32,48 > $O/${TAG}_e2e_inproc2.txt 2>&1; cat $O/${TAG}_e2e_inproc2.txt
echo "== e2e sweep 1 GPU"; timeout 400 python scripts/e2e_scaling.py --devices 1 --modes spin --threads 8,12,16,24,32 > $O/${TAG}_e2e_inproc1.txt 2>&1; cat $O/${TAG}_e2e_inproc1.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 3000 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; tail -c 3000 $O/${TAG}_bench_n2.json; tail -5 $O/${TAG}_bench_n2.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-400
ls -la $O | tail -20
