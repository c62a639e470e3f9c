#!/bin/bash
# r02o visit (1 GPU): who should fetch page-locked frames — copy engine, SMs, SMs with 256-byte L2 fetches; ncu of the
# row-form CRC32-C kernel
TAG=r02o
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
for k in ce sm sm256; do
  echo "== fetch by $k"; ACB200_FETCH=$k timeout 600 python scripts/e2e_scaling.py --devices 1 --register --modes spin --depths=-1 --threads 1,4,8,16 --seconds 0.8 2>&1 | grep -v "^registered"
done | tee $O/${TAG}_fetch_variants.txt
echo "== pytest (fetch by sm256)"; ACB200_FETCH=sm256 timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -x -k "fetch or transfer" 2>&1 | tail -3
echo "== ncu crc rows"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_crc32c_rows' -s 2 -c 1 \
    -o $O/${TAG}_crc_rows python scripts/prof_effects.py 256 > $O/${TAG}_ncu_crc_rows.log 2>&1; tail -2 $O/${TAG}_ncu_crc_rows.log
echo "== launch times crc"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_crc32c' -c 12 --csv --log-file $O/${TAG}_crc_launches.csv python scripts/prof_effects.py 256 > /dev/null 2>&1; tail -6 $O/${TAG}_crc_launches.csv | cut -c1-250
ls -la $O | tail -5
