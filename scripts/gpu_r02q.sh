#!/bin/bash
# r02q visit (8 GPUs): the drop-in call through ONE process driving 8 GPUs — pageable frames vs page-locked frames
# fetched by the devices; the reference arm on the same box
TAG=r02q
O=gpurun_out
mkdir -p $O
nproc > $O/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $O/${TAG}_gpu.txt; nvidia-smi -L | wc -l >> $O/${TAG}_gpu.txt
echo "== bench leg: e2e in one process, 8 GPUs"; timeout 600 python bench.py --leg e2e-inprocess --gpus 8 > $O/${TAG}_e2e_leg_n8.json 2> $O/${TAG}_e2e_leg_n8.err; tail -2 $O/${TAG}_e2e_leg_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_e2e_leg_n8.json").read().strip().splitlines()[-1])
for k in ("pageable", "registered"):
    r = d.get(k) or {}
    print(k, {x: r.get(x) for x in ("threads", "fetch_depth", "wait", "host_us_per_call", "ring_fingerprint")},
          "frames/s", round(r["calls"] / r["seconds"]) if r else None)
    print("   probes", r.get("probes"))
PY
echo "== sweep"; timeout 600 python scripts/e2e_scaling.py --devices 8 --register --modes spin,yield --depths=-1,4 --threads 32,48,64,96 --seconds 0.8 2>&1 | tee $O/${TAG}_e2e_inproc8.txt
echo "== reference arm"; timeout 600 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>/dev/null; cut -c1-200 $O/${TAG}_bench_reference.json
echo "== same leg, 1 GPU of this box"; timeout 600 python bench.py --leg e2e-inprocess --gpus 1 > $O/${TAG}_e2e_leg_n1.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_e2e_leg_n1.json").read().strip().splitlines()[-1])
for k in ("pageable", "registered"):
    r = d.get(k) or {}
    print(k, {x: r.get(x) for x in ("threads", "fetch_depth", "wait")}, "frames/s", round(r["calls"] / r["seconds"]) if r else None)
PY
