"""e2e drop-in call throughput vs caller threads x wait mode (tuning aid): where does the host path saturate?

    python scripts/e2e_scaling.py [--devices N] [--box]

ACB200_NN_PLAN=rows in the environment selects the round-1 row-granular transfer plan (A/B against the default
pixel-granular plan).  --devices N > 1: one process, N GPUs behind the C ABI (acb200_init_devices)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--box", action="store_true")
ap.add_argument("--threads", default="1,2,4,8,12,16,24,32,48,64")
ap.add_argument("--modes", default="spin,yield,block,hybrid")
ap.add_argument("--seconds", type=float, default=1.2)
ap.add_argument("--register", action="store_true", help="page-lock the frame ring (acb200_register_host_memory)")
ap.add_argument("--depths", default="", help="acb200_set_fetch_depth values to sweep (with --register), e.g. 0,2,4,-1")
a = ap.parse_args()
if a.devices > 1:
    assert acb.init_devices(list(range(a.devices))) == 0, acb.last_error()
else:
    assert acb.lib().acb200_init(0) == 0
H = bench.load_harness()
frames = bench.host_ring()
caps = acb.make_caps(bench.LEVEL, bench.MODE)
if a.register:
    import time
    t0 = time.perf_counter()
    rc = acb.lib().acb200_register_host_memory(frames.ctypes.data, frames.nbytes)
    print("registered %.2f GB of frames in %.2f s: rc %d %s" % (frames.nbytes / 1e9, time.perf_counter() - t0, rc,
                                                                 acb.last_error() if rc else ""), flush=True)
depths = [int(x) for x in a.depths.split(",")] if a.depths else [None]
fn = C.cast(acb.lib().ascii_convert_with_capabilities, C.c_void_p)
plan = os.environ.get("ACB200_NN_PLAN", "pixels")
print("cores %d  devices %d  nn plan %s  h2d %s" % (os.cpu_count(), a.devices, plan, os.environ.get("ACB200_H2D", "copy")))
scales = [(acb.SCALE_NN, "nn")] + ([(acb.SCALE_BOX, "box")] if a.box else [])
for scale, name in scales:
    acb.lib().acb200_set_default_scale(scale)
    for mode, depth in [(m, d) for m in a.modes.split(",") for d in depths]:
        acb.lib().acb200_set_sync_mode({"spin": 0, "block": 1, "hybrid": 2, "yield": 3}[mode], 30)
        if depth is not None:
            acb.lib().acb200_set_fetch_depth(depth)
            mode = "%s/d%d" % (mode, depth)
        for t in [int(x) for x in a.threads.split(",")]:
            ph = (C.c_uint64 * 5)()
            acb.lib().acb200_host_phase_stats(ph, 1)
            r = bench.run_callers(H, fn, frames, caps, t, a.seconds)
            acb.lib().acb200_host_phase_stats(ph, 1)
            fps = r["calls"] / r["seconds"]
            us = [ph[k] / max(1, ph[4]) / 1e3 for k in range(4)]
            print("%s %-9s threads %2d: %8.0f frames/s  %7.1f Gpix/s  %6.3f ms per call per thread  D2H %.1f GB/s"
                  "   per call: gather %.0f enqueue %.0f wait %.0f copy-out %.0f us" % (
                      name, mode, t, fps, fps * bench.MPIX / 1e3, 1e3 * t / fps,
                      fps * r["bytes"] / max(1, r["calls"]) / 1e9, *us), flush=True)
