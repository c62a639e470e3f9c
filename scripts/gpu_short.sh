#!/bin/bash
# short confirmation visit: parity tests, smoke, the headline workload, the bench line, launch list
TAG=${1:-short}; O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -3 $O/${TAG}_pytest.txt | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/${TAG}_smoke.txt
echo "== prof target"; timeout 200 python scripts/prof_target.py 256 2>&1 | tail -2 | tee $O/${TAG}_proftarget.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json | cut -c1-200
echo "== ncu launch list of the bench's resident loop"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render|k_stitch|k_dither|k_crc|k_color|k_grid|k_comp' \
    --csv --log-file $O/${TAG}_resident_launches.csv python bench.py --steps 3 --warmup 3 --ring 64 --resident-only > $O/${TAG}_ncu_list.log 2>&1; tail -1 $O/${TAG}_ncu_list.log | cut -c1-200
echo "== ncu full ws2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows_ws2 -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_ws2.log 2>&1; tail -1 $O/${TAG}_ncu_ws2.log
