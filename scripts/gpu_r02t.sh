#!/bin/bash
# r02t visit (1 GPU): smoke() with the device-side fetch check, GPU tests on the last commit
TAG=r02t
O=gpurun_out
mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -5 $O/${TAG}_pytest.txt | cut -c1-600
