"""e2e drop-in call throughput vs caller threads (tuning aid): where does the host path saturate?"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

assert acb.lib().acb200_init(0) == 0
H = bench.load_harness()
frames = bench.host_ring()
caps = acb.make_caps(bench.LEVEL, bench.MODE)
fn = C.cast(acb.lib().ascii_convert_with_capabilities, C.c_void_p)
for scale, name in ((acb.SCALE_NN, "nn"), (acb.SCALE_BOX, "box")):
    acb.lib().acb200_set_default_scale(scale)
    for t in (1, 2, 4, 8, 16, 32):
        r = bench.run_callers(H, fn, frames, caps, t, 1.5)
        fps = r["calls"] / r["seconds"]
        print("%s threads %2d: %8.0f frames/s  %6.3f ms per call per thread  H2D %.1f GB/s  D2H %.1f GB/s" % (
            name, t, fps, 1e3 * t / fps, fps * (2.2118 if name == "nn" else 24.8832) / 1e3, fps * r["bytes"] / r["calls"] / 1e9))
