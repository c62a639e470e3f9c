#!/bin/bash
# r02s visit (1 GPU): final tree — GPU tests, smoke, CRC32-C variants (cp.async ring / register ring / round-1 segments)
# with the PRMT+IMAD lookups, ncu of the rows kernel, bench line, parity sweep, sanitizer ladder
TAG=r02s
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -8 $O/${TAG}_pytest.txt | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
echo "== CRC32-C"; (timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/rows, cp.async ring: /'; ACB200_CRC_ROWS=regs timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/rows, register ring: /'; ACB200_CRC_KERNEL=segments timeout 300 python scripts/time_effects.py 2>&1 | grep frame_packets | sed 's/^/segments (round 1): /') | tee $O/${TAG}_crc_ab.txt
echo "== ncu crc rows"; timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_crc32c_rows' -s 2 -c 1 \
    -o $O/${TAG}_crc_rows python scripts/prof_effects.py 256 > $O/${TAG}_ncu_crc_rows.log 2>&1; tail -1 $O/${TAG}_ncu_crc_rows.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_crc32c' -c 9 --csv --log-file $O/${TAG}_crc_launches.csv python scripts/prof_effects.py 256 > /dev/null 2>&1; grep -o 'k_crc32c_[a-z]*\|"ns","[0-9]*"' $O/${TAG}_crc_launches.csv | paste - - | tail -3
echo "== bench N=1"; timeout 1200 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"])
for k in ("e2e", "e2e_pageable"):
    e = d.get(k) or {}
    print(k, {x: e.get(x) for x in ("value", "frames_per_s", "caller_threads", "wait", "input", "host_us_per_call", "ring_fingerprint", "bytes_identical_to_cpu_baseline")})
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), "frame_packets", d.get("frame_packets", {}).get("ms_per_batch"))
PY
tail -3 $O/${TAG}_bench_n1.err
echo "== fuzz 60 s"; timeout 400 python scripts/fuzz_parity.py 60 2>&1 | tail -4 | tee $O/${TAG}_fuzz_parity.txt
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | tee $O/${TAG}_compute_sanitizer.txt | tail -12
