#!/bin/bash
# One GPU-box visit: diagnosis ladder, parity tests, smoke, bench, ncu launch list + full capture of the top kernel.
# Usage: scripts/gpu_round.sh [tag].  Everything worth keeping is written under gpurun_out/ (merged back by gpurun).
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $O/${TAG}_gpu.txt
echo "== diag"; timeout 300 python scripts/diag.py > $O/${TAG}_diag.txt 2>&1; tail -40 $O/${TAG}_diag.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -60 $O/${TAG}_pytest.txt | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/${TAG}_smoke.txt
echo "== prof target (plain)"; timeout 300 python scripts/prof_target.py 2>&1 | tail -4 | tee $O/${TAG}_proftarget.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $O/${TAG}_bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_render_rows|k_stitch|k_dither' \
    --csv --log-file $O/${TAG}_launches.csv python scripts/prof_target.py 16 > $O/${TAG}_ncu_list.log 2>&1; tail -3 $O/${TAG}_ncu_list.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows -s 3 -c 1 \
    -o $O/${TAG}_render_rows python scripts/prof_target.py 16 > $O/${TAG}_ncu_full.log 2>&1; tail -3 $O/${TAG}_ncu_full.log
fi
ls -la $O | tail -20
