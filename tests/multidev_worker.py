"""Worker process of tests/test_gpu_multidev.py: ONE process, every visible GPU behind the C ABI.

Owns its process because acb200_init_devices() must come before any other call into the library.  Prints one JSON
line; the parent test asserts on it.  Checker = the compiled reference when oracle/_ref travelled, else the pinned port.
"""
import json
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import numpy as np  # noqa: E402

import ascii_chat_b200 as acb  # noqa: E402
import oracle_bind as ob  # noqa: E402


def main():
    want_devices = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    devs = list(range(want_devices)) if want_devices > 0 else None
    assert acb.init_devices(devs) == 0, acb.last_error()
    L = acb.lib()
    ndev = L.acb200_device_count()
    res = {"devices": ndev, "ordinals": [L.acb200_device_at(k) for k in range(ndev)]}
    conv = ob.ref_convert if ob.ref() is not None else ob.port_convert
    mixed = ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame
    grid = ob.ref_create_grid if ob.ref() is not None else ob.port_create_grid

    # 1. calling threads are spread over the pool; every device renders the reference's bytes
    img = ob.gen("noise", 1920, 1080, 5)
    exp = {(lv, md): conv(img, 160, 48, lv, md) for lv, md in ((2, 0), (3, 2), (0, 0), (3, 0))}
    seen, bad = {}, []

    def caller(i):
        for (lv, md), e in exp.items():
            got = acb.ascii_convert_with_capabilities(img, 160, 48, acb.make_caps(lv, md), False, False, "standard")
            if got != e:
                bad.append((i, lv, md))
        seen[i] = L.acb200_thread_device()
    ts = [threading.Thread(target=caller, args=(i,)) for i in range(max(4, 2 * ndev))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    res["thread_devices"] = sorted(set(seen.values()))
    res["convert_mismatches"] = bad

    # 2. resident sources sharded by slot; a viewer on ANY device composes sources that live on the others
    n, W, H = 8, 160, 48
    srcs = [ob.gen(("noise", "bars", "gradient")[i % 3], 1920, 1080, i) for i in range(n)]
    for i, s in enumerate(srcs):
        rc = acb.source_update(i, s) if i % 2 == 0 else acb.source_update_pinned(i, s)
        assert rc == 0, (i, acb.last_error())
    res["slot_devices"] = [L.acb200_source_device(i) for i in range(n)]
    exp_mixed = mixed(srcs, W, H, 2, 0, "standard", True)
    cells = [conv(s, 160, 48, 2, 0) for s in srcs]
    exp_grid = grid([c + b"\0" for c in cells], 320, 96)
    mixed_bad, grid_bad = [], []

    def viewer(k):
        assert L.acb200_bind_thread(k) == 0
        got = acb.mixed_frame(list(range(n)), W, H, acb.make_caps(2, 0, True), "standard")
        if got != exp_mixed:
            mixed_bad.append(k)
        g, size = acb.grid_frame(list(range(n)), 160, 48, acb.make_caps(2, 0), "standard", 320, 96)
        if (g, size) != (exp_grid[0], exp_grid[1]):
            grid_bad.append((k, size, exp_grid[1], None if g is None else len(g)))
    for rep in range(2):
        ts = [threading.Thread(target=viewer, args=(k,)) for k in range(ndev)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    res["mixed_mismatch_devices"] = mixed_bad
    res["grid_mismatch_devices"] = grid_bad
    res["grid_bytes"] = exp_grid[1]
    res["launches"] = acb.launch_count()
    for i in range(n):
        acb.source_clear(i)
    L.acb200_shutdown()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
