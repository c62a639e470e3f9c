#!/bin/bash
# r02c visit (2 GPUs): full GPU suite after the grid_frame fix, filtered streaming box, D2H A/B, NN-mode ncu, bench N=1/2.
TAG=r02c
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -40 $O/${TAG}_pytest.txt | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== filtered box"; timeout 300 python scripts/prof_filtered.py 64 2>&1 | tee $O/${TAG}_filtered_box.txt
echo "== configs"; timeout 600 python scripts/measure_configs.py > $O/${TAG}_configs.txt 2>&1; grep -E "'nn'" $O/${TAG}_configs.txt | cut -c1-200
echo "== D2H A/B (1 GPU)"; for m in zc ce zc ce; do ACB200_D2H=$m timeout 300 python scripts/e2e_scaling.py --devices 1 --modes spin --threads 1,8,16 --seconds 1.0 2>&1 | grep threads | sed "s/^/$m /"; done | tee $O/${TAG}_d2h_ab.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 1500 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; tail -c 1500 $O/${TAG}_bench_n2.json; tail -5 $O/${TAG}_bench_n2.err
echo "== ncu NN"; for c in noise flat; do timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows$' -s 2 -c 1 \
    -o $O/${TAG}_nn_$c python scripts/prof_target.py 256 $c > $O/${TAG}_ncu_nn_$c.log 2>&1; tail -2 $O/${TAG}_ncu_nn_$c.log; done
echo "== ncu filtered"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows_ws2$' -s 22 -c 1 \
    -o $O/${TAG}_filtered python scripts/prof_filtered.py 64 > $O/${TAG}_ncu_filtered.log 2>&1; tail -2 $O/${TAG}_ncu_filtered.log
ls -la $O | tail -12
