#!/bin/bash
# short GPU visit: correctness ladder first (bails out if it hangs), then the tuning sweep
O=gpurun_out; mkdir -p $O; TAG=${1:-q}
echo "== diag"; timeout 150 python scripts/diag.py > $O/${TAG}_diag.txt 2>&1; rc=$?; tail -15 $O/${TAG}_diag.txt
if [ $rc -ne 0 ]; then echo "diag failed rc=$rc, stopping"; exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-300
for v in "ACB200_DIRECT=1" "ACB200_DIRECT=0" $EXTRA_VARIANTS; do
  echo "== variant [$v]"
  env $v timeout 200 python scripts/prof_target.py 256 2>&1 | tail -2
done | tee $O/${TAG}_sweep.txt
