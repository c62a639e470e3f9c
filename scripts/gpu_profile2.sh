#!/bin/bash
O=gpurun_out; mkdir -p $O; TAG=${1:-p2}
echo "== configs"; timeout 600 python scripts/measure_configs.py 2>&1 | tail -18 | cut -c1-260
echo "== ncu launch list of the bench's resident loop"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render|k_stitch|k_dither' \
    --csv --log-file $O/${TAG}_resident_launches.csv python bench.py --steps 3 --warmup 3 --ring 64 --resident-only > $O/${TAG}_ncu_list.log 2>&1; tail -1 $O/${TAG}_ncu_list.log | cut -c1-200
