#!/bin/bash
# tuning sweep on the 4K half-block box workload (256 resident frames); prints GB/s of the row kernel
O=gpurun_out; mkdir -p $O
for v in "" "ACB200_TUNE_NOALIAS=1" "ACB200_PHASE_A_ONLY=1" $EXTRA_VARIANTS; do
  echo "== variant [$v]"
  env $v timeout 300 python scripts/prof_target.py 256 2>&1 | tail -2
done | tee $O/${1:-sweep}.txt
