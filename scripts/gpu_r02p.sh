#!/bin/bash
# r02p visit (1 GPU): row-form CRC32-C with the cp.async row ring; bench line with the two e2e legs (pageable / page-locked
# frames); parity sweep + sanitizer ladder on this tree
TAG=r02p
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -8 $O/${TAG}_pytest.txt | cut -c1-600
echo "== CRC32-C"; (timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/rows:     /'; ACB200_CRC_KERNEL=segments timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/segments: /') | tee $O/${TAG}_crc_ab.txt
echo "== ncu crc rows"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_crc32c_rows' -s 2 -c 1 \
    -o $O/${TAG}_crc_rows python scripts/prof_effects.py 256 > $O/${TAG}_ncu_crc_rows.log 2>&1; tail -1 $O/${TAG}_ncu_crc_rows.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_crc32c' -c 9 --csv --log-file $O/${TAG}_crc_launches.csv python scripts/prof_effects.py 256 > /dev/null 2>&1; grep -o 'k_crc32c_[a-z]*\|"ns","[0-9]*"' $O/${TAG}_crc_launches.csv | paste - - | tail -6
echo "== bench N=1"; timeout 1200 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench_n1.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step"): print(k, d[k])
print("roofline", d["roofline"]["frac"])
for k in ("e2e", "e2e_pageable"):
    e = d.get(k) or {}
    print(k, {x: e.get(x) for x in ("value", "frames_per_s", "caller_threads", "wait", "input", "host_us_per_call", "ring_fingerprint", "bytes_identical_to_cpu_baseline")})
    print("   probes", e.get("probes"))
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), "frame_packets", d.get("frame_packets"))
PY
tail -3 $O/${TAG}_bench_n1.err
echo "== fuzz 60 s"; timeout 400 python scripts/fuzz_parity.py 60 2>&1 | tail -4 | tee $O/${TAG}_fuzz_parity.txt
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt | tail -12
