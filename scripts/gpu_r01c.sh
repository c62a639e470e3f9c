#!/bin/bash
# GPU visit: A/B of the streaming reduce, parity, bench, ncu full capture (with source counters) of the headline kernel
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -15 $O/${TAG}_pytest.txt | cut -c1-600
for v in "ACB200_SLOW_REDUCE=0" "ACB200_SLOW_REDUCE=1" "ACB200_SLOW_REDUCE=0" "ACB200_SLOW_REDUCE=1" "ACB200_WS2_NOEMIT=1" $EXTRA_VARIANTS; do
  echo "== variant [$v]"
  env $v timeout 200 python scripts/prof_target.py 256 2>&1 | tail -2
done | tee $O/${TAG}_sweep.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/${TAG}_bench.json | cut -c1-1500
echo "== ncu full ws2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows_ws2 -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_full.log 2>&1; tail -2 $O/${TAG}_ncu_full.log
ls -la $O | tail -8
