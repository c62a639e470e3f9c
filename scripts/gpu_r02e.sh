#!/bin/bash
# r02e visit (1 GPU): digital rain parity, filtered box with the replicated table, bench, sanitizer on the new kernels
TAG=r02e
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== filtered box"; timeout 300 python scripts/prof_filtered.py 256 2>&1 | tee $O/${TAG}_filtered_box.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 400 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt
echo "== ncu filtered"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_render_rows_ws2$' -s 22 -c 1 \
    -o $O/${TAG}_filtered python scripts/prof_filtered.py 64 > $O/${TAG}_ncu_filtered.log 2>&1; tail -2 $O/${TAG}_ncu_filtered.log
ls -la $O | tail -8
