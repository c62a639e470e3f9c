#!/bin/bash
# GPU visit for the §8f rows (display path, colour filter, wire packaging): parity tests, smoke, both bench arms,
# launch list of the bench's resident loop, ncu captures of the new kernels.
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -40 $O/${TAG}_pytest.txt | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/${TAG}_smoke.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-400
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 | tee $O/${TAG}_bench.json | cut -c1-3000
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list (effects)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/${TAG}_effects_launches.csv python scripts/prof_effects.py > $O/${TAG}_ncu_list.log 2>&1; tail -2 $O/${TAG}_ncu_list.log
echo "== ncu full: k_color_filter"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_color_filter -s 2 -c 1 \
    -o $O/${TAG}_color_filter python scripts/prof_effects.py > $O/${TAG}_ncu_cf.log 2>&1; tail -2 $O/${TAG}_ncu_cf.log
echo "== ncu full: k_crc32c_chunks"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_crc32c_chunks -s 2 -c 1 \
    -o $O/${TAG}_crc32c python scripts/prof_effects.py > $O/${TAG}_ncu_crc.log 2>&1; tail -2 $O/${TAG}_ncu_crc.log
fi
ls -la $O | tail -20
