#!/bin/bash
# r02j visit (1 GPU): 384-thread NN variant; fresh ncu capture of the headline kernel for profiles/traffic.json
TAG=r02j
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-400
echo "== configs"; timeout 600 python scripts/measure_configs.py > $O/${TAG}_configs.txt 2>&1; grep -E "'nn'" $O/${TAG}_configs.txt | cut -c1-200
echo "== fuzz 100 s"; timeout 400 python scripts/fuzz_parity.py 100 2>&1 | tail -8 | tee $O/${TAG}_fuzz_parity.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 300 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== ncu ws2 (headline kernel, 64 frames)"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows_ws2$' -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_ws2.log 2>&1; tail -2 $O/${TAG}_ncu_ws2.log
echo "== ncu NN flat"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows$' -s 2 -c 1 \
    -o $O/${TAG}_nn_flat python scripts/prof_target.py 256 flat > $O/${TAG}_ncu_nn_flat.log 2>&1; tail -2 $O/${TAG}_ncu_nn_flat.log
ls -la $O | tail -6
