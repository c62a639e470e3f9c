#!/bin/bash
TAG=${1:-r01h}; O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -8 $O/${TAG}_pytest.txt | cut -c1-400
for v in "ACB200_CF_CTAS_PER_SM=0" "ACB200_CF_CTAS_PER_SM=8" "ACB200_CF_CTAS_PER_SM=16" "ACB200_CF_CTAS_PER_SM=32"; do
  echo "== [$v]"; env $v timeout 300 python scripts/time_effects.py 2>&1 | tail -2
done | tee $O/${TAG}_effects_sweep.txt
echo "== ncu full: k_crc32c_chunks"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_crc32c_chunks -s 2 -c 1 \
    -o $O/${TAG}_crc32c python scripts/prof_effects.py > $O/${TAG}_ncu_crc.log 2>&1; tail -1 $O/${TAG}_ncu_crc.log
