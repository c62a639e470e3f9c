#!/bin/bash
# A/B of experimental library builds against the default: build each variant with
#   ACB200_LIB_NAME=libexp_X.so ACB200_OBJ_PREFIX=exp_X_ ACB200_EXTRA_NVCC="-DSOME_EXPERIMENT_MACRO" python ascii-chat_b200/build.py --force
# (the macro guards whatever is being tried in csrc/), list the library names below, run under gpurun.
O=gpurun_out; TAG=${1:-exp}; mkdir -p $O
for round in 1 2; do
for v in libasciichat_b200.so libexp_A.so libexp_B.so libexp_C.so; do
  echo "== [$v]"; ACB200_LIB_NAME=$v timeout 200 python scripts/prof_target.py 256 2>&1 | grep "scale 1"
done; done | tee $O/${TAG}_sweep.txt
