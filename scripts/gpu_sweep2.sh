#!/bin/bash
O=gpurun_out; mkdir -p $O
for v in "ACB200_DIRECT=1" "ACB200_WS2_NOEMIT=1" $EXTRA_VARIANTS; do
  echo "== variant [$v]"; env $v timeout 200 python scripts/prof_target.py 256 2>&1 | tail -2 | head -1
done | tee $O/${1:-sweep2}.txt
