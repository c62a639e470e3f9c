#!/bin/bash
# r02y visit (1 GPU): the bench line on the last tree of the round
TAG=r02y
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02y_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "clocks", d["clocks"])
for k in ("e2e", "e2e_pageable"):
    e = d.get(k) or {}
    print(k, {x: e.get(x) for x in ("value", "frames_per_s", "caller_threads", "wait", "host_us_per_call", "bytes_identical_to_cpu_baseline")})
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), "nn", d.get("nn_noise"), d.get("nn_flat"), "frame_packets", d.get("frame_packets", {}).get("ms_per_batch"))
PY
tail -2 $O/${TAG}_bench_n1.err
