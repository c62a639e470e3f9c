#!/bin/bash
# r02l visit (2 GPUs): where does the host time of the drop-in call go, and does a yielding wait let it scale?
TAG=r02l
O=gpurun_out
mkdir -p $O
nvidia-smi -L; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA" 
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -5 $O/${TAG}_pytest.txt | cut -c1-400
echo "== e2e one process, 2 GPUs"; timeout 600 python scripts/e2e_scaling.py --devices 2 --modes spin,yield --threads 8,12,16,24,32,48,64 --seconds 1.0 2>&1 | tee $O/${TAG}_e2e_inproc2.txt
echo "== e2e one process, 2 GPUs, zero-copy input"; ACB200_H2D=zc timeout 600 python scripts/e2e_scaling.py --devices 2 --modes spin,yield --threads 12,16,24,32,48 --seconds 1.0 2>&1 | tee $O/${TAG}_e2e_inproc2_zc.txt
echo "== e2e 1 GPU"; timeout 600 python scripts/e2e_scaling.py --devices 1 --modes spin,yield --threads 1,8,12,16,24,32 --seconds 1.0 2>&1 | tee $O/${TAG}_e2e_inproc1.txt
