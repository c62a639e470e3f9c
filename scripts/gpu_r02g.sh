#!/bin/bash
# r02g visit (1 GPU): final-tree checks — GPU tests, smoke, bench with every leg, PCIe ceiling of the box
TAG=r02g
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -30 $O/${TAG}_pytest.txt | cut -c1-400
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== pcie peak"; timeout 300 python scripts/pcie_peak.py 2>&1 | tee $O/${TAG}_pcie_peak.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 1500 $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | tee $O/${TAG}_bench_reference.json | cut -c1-200
ls -la $O | tail -6
