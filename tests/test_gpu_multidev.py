"""One process, several GPUs behind the C ABI (acb200_init_devices): context pool per device, calling threads leased
round-robin, resident source slots sharded by client and read across GPUs over peer access, and the in-process grid
(acb200_grid_frame: cells rendered on the owning GPU, rows stored into the composing GPU's arena).

On a 1-GPU box the same code runs with a pool of one device; with >= 2 GPUs the cross-device paths are exercised
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multidev.py -m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_one_process_all_gpus_python():
    n = _gpus()
    assert n >= 1
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidev_worker.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["devices"] == n and res["ordinals"] == list(range(n))
    assert res["thread_devices"] == list(range(n)), res  # every device served a calling thread
    assert res["convert_mismatches"] == []
    assert res["slot_devices"] == [i % n for i in range(8)]  # client c lives on GPU c % N
    assert res["mixed_mismatch_devices"] == [] and res["grid_mismatch_devices"] == [], res
    assert res["launches"] > 0


def test_one_process_all_gpus_plain_c(ob):
    """tests/c/multi_gpu_drive.c: pthreads + the reference's entry point, nothing else"""
    exe = os.path.join(ROOT, "tests", "c", "multi_gpu_drive")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "c")], check=True)
    n = _gpus()
    r = subprocess.run([exe, str(max(8, 2 * n))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    f = dict(kv.split("=") for kv in r.stdout.strip().split())
    assert int(f["devices"]) == n and int(f["used"]) == n
    # same LCG noise frame as the C program (s0 = 12345, byte = s >> 24), rendered by the checker
    import numpy as np
    W, H = 640, 360
    state, buf = 12345, bytearray(W * H * 3)
    for i in range(W * H * 3):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        buf[i] = state >> 24
    img = np.frombuffer(bytes(buf), np.uint8).reshape(H, W, 3)
    conv = ob.ref_convert if ob.ref() is not None else ob.port_convert
    exp = conv(img, 100, 30, 3, 2, "standard")
    assert (int(f["len"]), f["fnv"]) == (len(exp), "%08x" % ob.fnv(exp))
