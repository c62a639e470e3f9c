#!/usr/bin/env python
"""Server per-client path (acb200_mixed_frame, resident sources) timed beside the compiled reference's
create_mixed_ascii_frame_for_client on the same inputs.  Measurement aid, not part of bench.py's contract."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_bind as ob  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402


def main():
    assert acb.lib().acb200_init(0) == 0
    res = []
    for (n, sw, sh, W, H, level, mode) in ((9, 1280, 720, 240, 67, 3, 2), (4, 1920, 1080, 203, 61, 3, 0),
                                           (1, 1920, 1080, 203, 61, 3, 2), (9, 640, 480, 120, 40, 2, 0)):
        srcs = [ob.gen(("noise", "gradient", "bars")[i % 3], sw, sh, i) for i in range(n)]
        for i, s in enumerate(srcs):
            assert acb.source_update(i, s) == 0
        caps = acb.make_caps(level, mode, True)
        slots = list(range(n))
        got = acb.mixed_frame(slots, W, H, caps, "standard")
        exp = (ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame)(srcs, W, H, level, mode, "standard", True)
        ok = got == exp
        for _ in range(20):
            acb.mixed_frame(slots, W, H, caps, "standard")
        t0 = time.perf_counter()
        K = 200
        for _ in range(K):
            acb.mixed_frame(slots, W, H, caps, "standard")
        one = (time.perf_counter() - t0) / K
        # n render threads (one per receiving client) + the receive side re-uploading every source at 60 Hz pace-free
        T, per = n, 100
        def render():
            for _ in range(per):
                acb.mixed_frame(slots, W, H, caps, "standard")
        ts = [threading.Thread(target=render) for _ in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        multi = (time.perf_counter() - t0) / (T * per)
        t0 = time.perf_counter()
        for _ in range(20):
            for i, s in enumerate(srcs):
                acb.source_update(i, s)
        upd = (time.perf_counter() - t0) / (20 * n)
        R = 10
        t0 = time.perf_counter()
        for _ in range(R):
            (ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame)(srcs, W, H, level, mode, "standard", True)
        ref = (time.perf_counter() - t0) / R
        res.append(dict(clients=n, source="%dx%d" % (sw, sh), terminal="%dx%d" % (W, H), level=level, mode=mode,
                        identical=ok, bytes=got[1], b200_ms_per_frame_1thread=one * 1e3,
                        b200_ms_per_frame_nthreads=multi * 1e3, b200_source_update_ms=upd * 1e3,
                        reference_cpu_ms_per_frame_1thread=ref * 1e3,
                        ref_kind="reference" if ob.ref() is not None else "port"))
        print(json.dumps(res[-1]), flush=True)
        for i in range(n):
            acb.source_clear(i)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "server_path.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
