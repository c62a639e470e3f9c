#!/bin/bash
TAG=${1:-r01j}; O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -12 $O/${TAG}_pytest.txt | cut -c1-500
echo "== C4 at 1"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29519 \
   scripts/measure_c4.py 2>&1 | tail -1 | tee $O/${TAG}_c4_n1.json | cut -c1-1200
for tool in memcheck racecheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize.py > $O/${TAG}_san_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|mismatches|Error|hazard" $O/${TAG}_san_$tool.txt | head -8
done
