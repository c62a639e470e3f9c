#!/bin/bash
# final visit of a round: parity, smoke, both bench arms, effects timings, launch lists, ncu full captures of the three kernels
TAG=${1:-final}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -8 $O/${TAG}_pytest.txt | cut -c1-500
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== effects"; timeout 300 python scripts/time_effects.py 2>&1 | tail -2 | tee $O/${TAG}_effects.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-200
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json | cut -c1-200
echo "== ncu launch list of the bench's resident loop"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render|k_stitch|k_dither|k_crc|k_color|k_grid|k_comp' \
    --csv --log-file $O/${TAG}_resident_launches.csv python bench.py --steps 3 --warmup 3 --ring 64 --resident-only > $O/${TAG}_ncu_list.log 2>&1; tail -1 $O/${TAG}_ncu_list.log | cut -c1-200
echo "== ncu launch list (effects)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/${TAG}_effects_launches.csv python scripts/prof_effects.py > $O/${TAG}_ncu_list2.log 2>&1; tail -1 $O/${TAG}_ncu_list2.log
echo "== ncu full ws2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows_ws2 -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_ws2.log 2>&1; tail -1 $O/${TAG}_ncu_ws2.log
echo "== ncu full: k_color_filter"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_filter -s 2 -c 1 \
    -o $O/${TAG}_color_filter python scripts/prof_effects.py > $O/${TAG}_ncu_cf.log 2>&1; tail -1 $O/${TAG}_ncu_cf.log
echo "== ncu full: k_crc32c_chunks"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_crc32c_chunks -s 2 -c 1 \
    -o $O/${TAG}_crc32c python scripts/prof_effects.py > $O/${TAG}_ncu_crc.log 2>&1; tail -1 $O/${TAG}_ncu_crc.log
