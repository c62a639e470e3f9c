#!/bin/bash
# r02u visit (1 GPU): CRC32-C on small batches (one-frame packet path) row form vs segment form; long parity sweep
TAG=r02u
O=gpurun_out
mkdir -p $O
echo "== small-batch CRC"; (timeout 300 python scripts/time_crc_small.py; ACB200_CRC_KERNEL=segments timeout 300 python scripts/time_crc_small.py) 2>&1 | tee $O/${TAG}_crc_small.txt
echo "== fuzz 150 s"; timeout 600 python scripts/fuzz_parity.py 150 2>&1 | tail -4 | tee $O/${TAG}_fuzz_parity.txt
