#!/bin/bash
# sanitizer ladder (incl. the §8f kernels), full bench both arms, launch list of the bench's resident loop
TAG=${1:-r01e}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== plain ladder"; timeout 300 python scripts/sanitize.py 2>&1 | tail -3
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize.py > $O/${TAG}_san_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|mismatches|Error|hazard" $O/${TAG}_san_$tool.txt | head -8
done
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json | cut -c1-300
echo "== ncu launch list of the bench's resident loop"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render|k_stitch|k_dither|k_crc|k_color' \
    --csv --log-file $O/${TAG}_resident_launches.csv python bench.py --steps 3 --warmup 3 --ring 64 --resident-only > $O/${TAG}_ncu_list.log 2>&1; tail -1 $O/${TAG}_ncu_list.log | cut -c1-200
