#!/bin/bash
# r02d visit (1 GPU): look-back record / scan-fold changes, fuzz sweep and sanitizer ladder on the round-2 tree, bench.
TAG=r02d
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-400
echo "== configs"; timeout 600 python scripts/measure_configs.py > $O/${TAG}_configs.txt 2>&1; grep -E "C3|truecolor fg" $O/${TAG}_configs.txt | cut -c1-200
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 600 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== fuzz 150 s"; timeout 400 python scripts/fuzz_parity.py 150 2>&1 | tail -8 | tee $O/${TAG}_fuzz_parity.txt
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt
echo "== ncu NN flat"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows$' -s 2 -c 1 \
    -o $O/${TAG}_nn_flat python scripts/prof_target.py 256 flat > $O/${TAG}_ncu_nn_flat.log 2>&1; tail -2 $O/${TAG}_ncu_nn_flat.log
echo "== ncu launch list (bench resident step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --resident-only --steps 2 --warmup 1 > $O/${TAG}_launches.log 2>&1; tail -2 $O/${TAG}_launches.log
ls -la $O | tail -12
