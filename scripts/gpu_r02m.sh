#!/bin/bash
# r02m visit (2 GPUs): frames in registered (page-locked) memory — the GPU fetches the sampled rows itself; fetch depth x
# wait mode x caller threads; A/B of the host gather's software prefetch
TAG=r02m
O=gpurun_out
mkdir -p $O
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; free -g | head -2
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -x 2>&1 > $O/${TAG}_pytest.txt; tail -15 $O/${TAG}_pytest.txt | cut -c1-600
echo "== registered ring, 2 GPUs"; timeout 900 python scripts/e2e_scaling.py --devices 2 --register --modes yield --depths 0,2,4,8,-1 --threads 8,16,24,32,48 --seconds 0.8 2>&1 | tee $O/${TAG}_e2e_registered2.txt
timeout 600 python scripts/e2e_scaling.py --devices 2 --register --modes spin --depths 0,4,-1 --threads 16,32 --seconds 0.8 2>&1 | tee -a $O/${TAG}_e2e_registered2.txt
echo "== registered ring, 1 GPU"; timeout 600 python scripts/e2e_scaling.py --devices 1 --register --modes yield --depths 0,2,4,-1 --threads 4,8,16,24 --seconds 0.8 2>&1 | tee $O/${TAG}_e2e_registered1.txt
echo "== gather prefetch A/B (pageable ring, 1 GPU)"
for ah in 0 2; do echo "ACB200_GATHER_AHEAD=$ah"; ACB200_GATHER_AHEAD=$ah timeout 600 python scripts/e2e_scaling.py --devices 1 --modes spin --threads 1,8,16 --seconds 0.8 2>&1; done | tee $O/${TAG}_gather_prefetch_ab.txt
