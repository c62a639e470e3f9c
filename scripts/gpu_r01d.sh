#!/bin/bash
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -15 $O/${TAG}_pytest.txt | cut -c1-600
for v in "ACB200_X=0" "ACB200_SLOW_REDUCE=1" "ACB200_X=0" "ACB200_WS2_NOEMIT=1" $EXTRA_VARIANTS; do
  echo "== variant [$v]"
  env $v timeout 200 python scripts/prof_target.py 256 2>&1 | tail -2
done | tee $O/${TAG}_sweep.txt
echo "== configs"; timeout 600 python scripts/measure_configs.py 2>&1 | tail -18 | cut -c1-260 | tee $O/${TAG}_configs.txt
echo "== ncu full ws2"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_rows_ws2 -s 3 -c 1 \
    -o $O/${TAG}_ws2 python scripts/prof_target.py 64 > $O/${TAG}_ncu_full.log 2>&1; tail -2 $O/${TAG}_ncu_full.log
