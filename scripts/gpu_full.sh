#!/bin/bash
# full evidence visit: tests, smoke, bench, ncu launch list + full capture of the headline kernel
TAG=${1:-full}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
echo "== diag"; timeout 150 python scripts/diag.py > $O/${TAG}_diag.txt 2>&1; rc=$?; tail -3 $O/${TAG}_diag.txt
if [ $rc -ne 0 ]; then echo "diag failed rc=$rc, stopping"; exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -5 $O/${TAG}_pytest.txt | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu launch list (bench command, 64-frame ring)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render_rows|k_stitch|k_dither|k_text|k_comp|k_resize' \
    --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --ring 64 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; tail -2 $O/${TAG}_ncu_list.log | cut -c1-200
echo "== ncu full (headline kernel, 64 frames)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows_ws2 -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_full.log 2>&1; tail -2 $O/${TAG}_ncu_full.log
ls -la $O | grep ${TAG}
