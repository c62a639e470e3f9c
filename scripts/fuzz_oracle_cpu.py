"""CPU-only: randomised port-vs-compiled-reference sweep (pins oracle/ascii_oracle.c harder than the fixed tests).
Usage: fuzz_oracle_cpu.py [seconds] [seed].  Needs oracle/_ref (build container only)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import oracle_bind as ob  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
assert ob.ref() is not None
PATS = ("noise", "bars", "gradient", "grey", "solid")
PALS = ("standard", "blocks", "digital", "minimal", "cool")
counts = {"display": 0, "mixed": 0, "grid": 0, "filter": 0, "crc": 0, "dither": 0, "rain": 0, "rainbow": 0, "box": 0}
t_end = time.time() + budget
i = 0
while time.time() < t_end:
    i += 1
    fam = i % 12
    if fam == 8:  # the three dithered leaf printers
        w, h = int(rng.integers(1, 140)), int(rng.integers(1, 70))
        img = ob.gen(PATS[int(rng.integers(0, 5))], w, h, i)
        pal, v = PALS[int(rng.integers(0, 5))], int(rng.integers(0, 3))
        assert ob.ref_print_dither(img, pal, v) == ob.port_print_dither(img, pal, v), (w, h, pal, v)
        counts["dither"] += 1
    elif fam == 9:  # digital rain over three frames (state), random grid sizes around the frame's
        cols, rows = int(rng.integers(1, 160)), int(rng.integers(1, 50))
        level, mode, pal = int(rng.integers(0, 4)), int(rng.integers(0, 3)), PALS[int(rng.integers(0, 5))]
        filt = int(rng.integers(0, 13))
        gc, gr = max(1, cols + int(rng.integers(-4, 5))), max(1, rows + int(rng.integers(-4, 5)))
        a, b = ob.RefRain(gc, gr, filt), ob.PortRain(gc, gr, filt)
        for k in range(3):
            s_ = ob.port_convert(ob.gen(PATS[int(rng.integers(0, 5))], int(rng.integers(8, 300)), int(rng.integers(8, 200)), i + k),
                                 cols, rows, level, mode, pal)
            if rng.random() < 0.2:
                cut = int(rng.integers(0, len(s_) + 1))
                s_ = s_[:cut].replace(b"\0", b"")
            dt = float(rng.random() * 0.2)
            assert a.apply(s_, dt) == b.apply(s_, dt), (cols, rows, level, mode, pal, filt, gc, gr, k)
        a.close()
        b.close()
        counts["rain"] += 1
    elif fam == 10:  # rainbow replace on rendered strings (never ending exactly on a colour code)
        s_ = ob.port_convert(ob.gen(PATS[int(rng.integers(0, 5))], int(rng.integers(8, 300)), int(rng.integers(8, 200)), i),
                             int(rng.integers(1, 160)), int(rng.integers(1, 50)), int(rng.integers(0, 4)), int(rng.integers(0, 3)),
                             PALS[int(rng.integers(0, 5))]) + b"."
        t = float(rng.random() * 30)
        assert ob.ref_rainbow_replace(s_, t) == ob.port_rainbow_replace(s_, t), (len(s_), t)
        counts["rainbow"] += 1
    elif fam == 11:  # the box-mode checker and the fast box filter
        W, H = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        img = ob.gen(PATS[int(rng.integers(0, 5))], W, H, i)
        c, r = int(rng.integers(1, 200)), int(rng.integers(1, 60))
        level, mode = int(rng.integers(0, 4)), int(rng.integers(0, 3))
        assert ob.ref_box_convert(img, c, r, level, mode) == ob.port_convert(img, c, r, level, mode, scale=ob.SCALE_BOX), \
            (W, H, c, r, level, mode)
        import ctypes as C
        u8p = C.POINTER(C.c_uint8)
        out = np.empty((r, c, 3), np.uint8)
        ob.port().orc_resize_box_fast(img.ctypes.data_as(u8p), W, H, out.ctypes.data_as(u8p), c, r)
        assert np.array_equal(out, ob.port_resize(img, c, r, scale=ob.SCALE_BOX)), (W, H, c, r)
        counts["box"] += 1
    elif fam < 4:
        W, H = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        img = ob.gen(PATS[int(rng.integers(0, 5))], W, H, i)
        if rng.random() < 0.3:
            x0 = int(rng.integers(0, W))
            img[:, x0:x0 + int(rng.integers(1, 2 + W // 2))] = 0
        kw = dict(cols=int(rng.integers(1, 200)), rows=int(rng.integers(1, 70)), level=int(rng.integers(-1, 4)),
                  mode=int(rng.integers(0, 3)), palette=PALS[int(rng.integers(0, 5))], aspect=bool(rng.integers(0, 2)),
                  stretch=bool(rng.integers(0, 2)), pad=bool(rng.integers(0, 2)), flip_x=bool(rng.integers(0, 2)),
                  flip_y=bool(rng.integers(0, 2)), color_filter=int(rng.integers(-1, 14)), time_s=float(rng.random() * 30))
        assert ob.ref_display_convert(img, **kw) == ob.port_display_convert(img, **kw), (img.shape, kw)
        counts["display"] += 1
    elif fam == 4:
        n = int(rng.integers(1, 12))
        srcs = [None if rng.random() < 0.15 else ob.gen(PATS[int(rng.integers(0, 4))], int(rng.integers(20, 400)),
                                                       int(rng.integers(16, 300)), k) for k in range(n)]
        W, H = int(rng.integers(20, 220)), int(rng.integers(8, 60))
        level, mode, pad = int(rng.integers(0, 4)), int(rng.integers(0, 3)), bool(rng.integers(0, 2))
        if ob.composite_degenerate(srcs, W, H):
            continue
        assert ob.ref_mixed_frame(srcs, W, H, level, mode, "standard", pad) == \
            ob.port_mixed_frame(srcs, W, H, level, mode, "standard", pad), (n, W, H, level, mode, pad)
        counts["mixed"] += 1
    elif fam == 5:
        n = int(rng.integers(1, 12))
        level, mode = int(rng.integers(0, 4)), int(rng.choice([0, 2]))
        cols, rows = int(rng.integers(4, 60)), int(rng.integers(2, 20))
        srcs = [ob.port_convert(ob.gen(PATS[k % 4], 96, 64, k), cols, rows, level, mode) for k in range(n)]
        W, H = int(rng.integers(10, 260)), int(rng.integers(3, 80))
        a, b = ob.ref_create_grid(srcs, W, H), ob.port_create_grid(srcs, W, H)
        canvas = W * H + H
        assert a == b or (len(a[0]) > canvas and b[0] == a[0][:canvas]), (n, level, mode, cols, rows, W, H)
        counts["grid"] += 1
    elif fam == 6:
        W, H = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        img = ob.gen(PATS[int(rng.integers(0, 5))], W, H, i)
        f, t = int(rng.integers(-1, 14)), float(rng.random() * 20)
        a, b = ob.ref_color_filter(img, f, t), ob.port_color_filter(img, f, t)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), (img.shape, f, t)
        counts["filter"] += 1
    else:
        L = int(rng.integers(0, 300000))
        d = rng.integers(0, 256, L, dtype=np.uint8).tobytes()
        assert ob.ref().ref_oracle_crc32(d, L) == ob.port().orc_crc32c(d, L), L
        assert ob.ref_packet_header(d, 200, 60) == ob.port_packet_header(d, 200, 60)
        counts["crc"] += 1
print("port == compiled reference (%.0f s): %s" % (budget, ", ".join("%s %d" % kv for kv in counts.items())))
