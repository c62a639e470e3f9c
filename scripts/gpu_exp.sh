#!/bin/bash
# A/B of experimental library variants (ascii-chat_b200/lib/libexp_*.so, built with -DACB_EXP_*) against the default
O=gpurun_out; TAG=${1:-exp}; mkdir -p $O
for round in 1 2; do
for v in libasciichat_b200.so libexp_A.so libexp_B.so libexp_C.so; do
  echo "== [$v]"; ACB200_LIB_NAME=$v timeout 200 python scripts/prof_target.py 256 2>&1 | grep "scale 1"
done; done | tee $O/${TAG}_sweep.txt
