#!/bin/bash
# r02v visit (1 GPU): last commit — GPU tests (CRC through both forms), smoke, sanitizer ladder
TAG=r02v
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -5 $O/${TAG}_pytest.txt | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${TAG}_smoke.txt
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt | tail -12
