#!/bin/bash
# r02a visit (1 GPU): parity on the new host path / persistent NN kernel / multi-device pool of one, e2e sweeps,
# TMA A/B, NN-mode and TMA ncu captures.
TAG=r02a
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== e2e sweep (pixel plan)"; timeout 600 python scripts/e2e_scaling.py > $O/${TAG}_e2e_sweep_pixels.txt 2>&1; cat $O/${TAG}_e2e_sweep_pixels.txt
echo "== e2e sweep (row plan)"; ACB200_NN_PLAN=rows timeout 300 python scripts/e2e_scaling.py --modes hybrid --threads 1,8,16,32 > $O/${TAG}_e2e_sweep_rows.txt 2>&1; cat $O/${TAG}_e2e_sweep_rows.txt
echo "== configs"; timeout 600 python scripts/measure_configs.py > $O/${TAG}_configs.txt 2>&1; grep -E "C3|truecolor fg" $O/${TAG}_configs.txt | cut -c1-260
echo "== TMA A/B"; for k in split tma split tma; do ACB200_BOX_KERNEL=$k timeout 300 python bench.py --resident-only --steps 20 --warmup 5 2>&1 | tail -1 | sed "s/^/$k /"; done | tee $O/${TAG}_tma_ab.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json | cut -c1-1500
echo "== ncu NN"; for c in noise flat; do timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows$' -s 2 -c 1 \
    -o $O/${TAG}_nn_$c python scripts/prof_target.py 256 $c > $O/${TAG}_ncu_nn_$c.log 2>&1; tail -2 $O/${TAG}_ncu_nn_$c.log; done
echo "== ncu TMA"; ACB200_BOX_KERNEL=tma timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows_ws$' -s 3 -c 1 \
    -o $O/${TAG}_tma python scripts/prof_target.py 64 > $O/${TAG}_ncu_tma.log 2>&1; tail -2 $O/${TAG}_ncu_tma.log
ls -la $O | tail -20
