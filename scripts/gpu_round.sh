#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list.  Usage: scripts/gpu_round.sh [tag]
# Everything worth keeping is written under gpurun_out/ (merged back by gpurun).
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench.json
