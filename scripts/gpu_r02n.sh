#!/bin/bash
# r02n visit (2 GPUs): copy-engine fetch of page-locked frames (fetch depth x wait mode x callers), row-form CRC32-C A/B
TAG=r02n
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -x 2>&1 > $O/${TAG}_pytest.txt; tail -15 $O/${TAG}_pytest.txt | cut -c1-700
echo "== CRC32-C: rows (default) vs segments (round 1)"
(timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/rows:     /'; ACB200_CRC_KERNEL=segments timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/segments: /') | tee $O/${TAG}_crc_ab.txt
echo "== registered ring, 2 GPUs"; timeout 900 python scripts/e2e_scaling.py --devices 2 --register --modes spin,yield --depths 0,2,4,-1 --threads 8,16,24,32 --seconds 0.8 2>&1 | tee $O/${TAG}_e2e_registered2.txt
echo "== registered ring, 1 GPU"; timeout 600 python scripts/e2e_scaling.py --devices 1 --register --modes spin --depths 0,2,4,-1 --threads 1,4,8,16,24 --seconds 0.8 2>&1 | tee $O/${TAG}_e2e_registered1.txt
