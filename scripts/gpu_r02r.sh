#!/bin/bash
# r02r visit (2 GPUs): the bench line at N=2 as the driver launches it (torchrun), on the tree with the two e2e legs
TAG=r02r
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; tail -3 $O/${TAG}_bench_n2.err | cut -c1-300
python - <<'PY'
import json
lines = [l for l in open("gpurun_out/r02r_bench_n2.json").read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("value", d["value"], "n_gpus", d["n_gpus"], "frac", d["roofline"]["frac"])
for k in ("e2e", "e2e_pageable", "e2e_per_rank_processes"):
    e = d.get(k) or {}
    print(k, {x: e.get(x) for x in ("value", "frames_per_s", "caller_threads", "wait", "input", "structure", "host_us_per_call", "ring_fingerprint", "bytes_identical_to_cpu_baseline", "caller_threads_per_gpu")})
print("c4", json.dumps(d.get("c4"))[:600])
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"))
PY
echo "== reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
