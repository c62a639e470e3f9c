"""Time-bounded randomised parity sweep on the GPU: every public entry against the compiled reference (oracle/_ref)
where it travelled, else the pinned port.  Usage: fuzz_parity.py [seconds] [seed].  Prints one summary line per
family; exits non-zero on the first mismatch (after printing the failing case)."""
import struct
import sys
import time
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402
import oracle_bind as ob  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261017
rng = np.random.default_rng(seed)
assert acb.lib().acb200_init(0) == 0
use_ref = ob.ref() is not None
want_display = ob.ref_display_convert if use_ref else ob.port_display_convert
want_mixed = ob.ref_mixed_frame if use_ref else ob.port_mixed_frame
want_grid = ob.ref_create_grid if use_ref else ob.port_create_grid
want_hdr = ob.ref_packet_header if use_ref else ob.port_packet_header
PATS = ("noise", "bars", "gradient", "grey", "solid")
PALS = ("standard", "blocks", "digital", "minimal", "cool")
want_dither = ob.ref_print_dither if use_ref else ob.port_print_dither
want_rainbow = ob.ref_rainbow_replace if use_ref else ob.port_rainbow_replace
want_conv = ob.ref_convert if use_ref else ob.port_convert
mk_rain = ob.RefRain if use_ref else ob.PortRain
counts = {"display": 0, "mixed+packet": 0, "grid": 0, "filter": 0, "dither": 0, "rainbow": 0, "rain": 0, "grid_frame": 0,
          "box": 0}
_trace = open(os.environ["FUZZ_TRACE"], "w") if os.environ.get("FUZZ_TRACE") else None


def trace(*a):  # last line of the trace file = the case that was running if the process dies
    if _trace:
        _trace.seek(0)
        _trace.truncate()
        _trace.write(repr(a) + "\n")
        _trace.flush()

t_end = time.time() + budget


def fail(kind, case):
    print("MISMATCH", kind, case)
    sys.exit(1)


def rnd_img(i):
    W, H = int(rng.integers(1, 700)), int(rng.integers(1, 500))
    img = ob.gen(PATS[int(rng.integers(0, 5))], W, H, i)
    if rng.random() < 0.3:  # black stripes: transparent half-block runs, REP runs
        x0 = int(rng.integers(0, W))
        img[:, x0:x0 + int(rng.integers(1, 2 + W // 2))] = 0
    return img


i = 0
while time.time() < t_end:
    i += 1
    fam = i % 12
    if fam == 8:  # the three dithered leaf printers (foreground.c:650-846)
        w, h = int(rng.integers(1, 120)), int(rng.integers(1, 60))
        img = ob.gen(PATS[int(rng.integers(0, 5))], w, h, i)
        pal, variant = PALS[int(rng.integers(0, 5))], int(rng.integers(0, 3))
        trace("dither", i, w, h, pal, variant)
        if acb.image_print_16color_dithered(img, pal, None if variant == 2 else variant == 0) != want_dither(img, pal, variant):
            fail("dither", (w, h, pal, variant))
        counts["dither"] += 1
    elif fam == 9:  # rainbow_replace_ansi_colors + digital rain on rendered strings (two frames: the carried state)
        img = rnd_img(i)
        cols, rows, level, mode = int(rng.integers(1, 200)), int(rng.integers(1, 60)), int(rng.integers(0, 4)), int(rng.integers(0, 3))
        pal = PALS[int(rng.integers(0, 5))]
        s1 = ob.port_convert(img, cols, rows, level, mode, pal)
        t = float(rng.random() * 20)
        trace("rainbow", i, img.shape, cols, rows, level, mode, pal, t)
        if acb.rainbow_replace_ansi_colors(s1 + b"!", t) != want_rainbow(s1 + b"!", t):
            fail("rainbow", (img.shape, cols, rows, level, mode, pal, t))
        counts["rainbow"] += 1
        filt = int(rng.integers(0, 13))
        gc, gr = max(1, cols + int(rng.integers(-3, 4))), max(1, rows + int(rng.integers(-3, 4)))
        a, b = acb.DigitalRain(gc, gr, filt), mk_rain(gc, gr, filt)
        for k in range(2):
            dt = float(rng.random() * 0.1)
            if a.apply(s1, dt) != b.apply(s1, dt):
                fail("rain", (img.shape, cols, rows, level, mode, pal, filt, gc, gr, k))
        a.close()
        b.close()
        counts["rain"] += 1
    elif fam == 10:  # acb200_grid_frame: resident slots -> cells -> ascii_create_grid (host.c:664-717)
        n = int(rng.integers(1, 10))
        srcs = [ob.gen(PATS[int(rng.integers(0, 4))], int(rng.integers(20, 400)), int(rng.integers(16, 300)), k) for k in range(n)]
        cw, ch = int(rng.integers(4, 100)), int(rng.integers(2, 30))
        W, H = int(rng.integers(10, 260)), int(rng.integers(3, 80))
        level, mode = int(rng.integers(0, 4)), int(rng.choice([0, 2]))
        trace("grid_frame", i, n, cw, ch, W, H, level, mode)
        for k, s_ in enumerate(srcs):
            acb.source_update(k, s_)
        got = acb.grid_frame(list(range(n)), cw, ch, acb.make_caps(level, mode), "standard", W, H)
        exp = want_grid([want_conv(s_, cw, ch, level, mode) + b"\0" for s_ in srcs], W, H)
        canvas = W * H + H
        if got != (exp[0], exp[1]) and not (exp[0] is not None and len(exp[0]) > canvas and got[0] == exp[0][:canvas]):
            fail("grid_frame", (n, cw, ch, W, H, level, mode, [s_.shape for s_ in srcs]))
        for k in range(n):
            acb.source_clear(k)
        counts["grid_frame"] += 1
    elif fam == 11:  # box mode: the compiled reference's printer on the box-filtered image; filters on the streaming path
        W16, H = 16 * int(rng.integers(2, 60)), int(rng.integers(8, 400))
        img = ob.gen(PATS[int(rng.integers(0, 5))], W16, H, i)
        c, r = int(rng.integers(1, W16 + 1)), int(rng.integers(1, 60))
        level, mode, pal = int(rng.integers(0, 4)), int(rng.integers(0, 3)), PALS[int(rng.integers(0, 5))]
        filt, fx, fy = int(rng.integers(0, 13)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        trace("box", i, W16, H, c, r, level, mode, pal, filt, fx, fy)
        cfg = acb.make_cfg(W16, H, c, r * 2 if mode == 2 else r, level, mode, pal, scale=acb.SCALE_BOX, flip_x=fx, flip_y=fy,
                           color_filter=filt, filter_time=1.0)
        got = acb.render_batch_host(cfg, [img])[0]
        exp = ob.port_display_convert(img, c, r, level, mode, pal, flip_x=fx, flip_y=fy, color_filter=filt, time_s=1.0,
                                      scale=ob.SCALE_BOX)
        if got != exp:
            fail("box", (W16, H, c, r, level, mode, pal, filt, fx, fy))
        if use_ref and filt == 0 and not fx and not fy and got != ob.ref_box_convert(img, c, r, level, mode, pal):
            fail("box/ref printer", (W16, H, c, r, level, mode, pal))
        counts["box"] += 1
    elif fam < 5:
        img = rnd_img(i)
        kw = dict(cols=int(rng.integers(1, 260)), rows=int(rng.integers(1, 90)), level=int(rng.integers(-1, 4)),
                  mode=int(rng.integers(0, 3)), palette=PALS[int(rng.integers(0, 5))], aspect=bool(rng.integers(0, 2)),
                  stretch=bool(rng.integers(0, 2)), pad=bool(rng.integers(0, 2)), flip_x=bool(rng.integers(0, 2)),
                  flip_y=bool(rng.integers(0, 2)), color_filter=int(rng.integers(-1, 14)), time_s=float(rng.random() * 30))
        trace("display", i, img.shape, kw)
        got = acb.display_convert(img, kw["cols"], kw["rows"], acb.make_caps(kw["level"], kw["mode"], kw["pad"]),
                                  kw["aspect"], kw["stretch"], kw["palette"], kw["flip_x"], kw["flip_y"],
                                  kw["color_filter"], kw["time_s"])
        if got != want_display(img, **kw):
            fail("display", (img.shape, kw))
        counts["display"] += 1
    elif fam == 5:
        n = int(rng.integers(1, 11))
        srcs = [None if rng.random() < 0.15 else ob.gen(PATS[int(rng.integers(0, 4))], int(rng.integers(20, 500)),
                                                       int(rng.integers(16, 400)), k) for k in range(n)]
        W, H = int(rng.integers(20, 240)), int(rng.integers(8, 70))
        level, mode, pad = int(rng.integers(0, 4)), int(rng.integers(0, 3)), bool(rng.integers(0, 2))
        if ob.composite_degenerate(srcs, W, H):
            continue
        trace("mixed", i, n, W, H, level, mode, pad, [None if s is None else s.shape for s in srcs])
        for k, s in enumerate(srcs):
            if s is None:
                acb.source_clear(k)
            elif k % 2:
                acb.source_update_wire(k, struct.pack(">II", s.shape[1], s.shape[0]) + s.tobytes())
            else:
                acb.source_update(k, s)
        caps = acb.make_caps(level, mode, pad)
        exp = want_mixed(srcs, W, H, level, mode, "standard", pad)
        if acb.mixed_frame(list(range(n)), W, H, caps, "standard") != exp:
            fail("mixed", (n, W, H, level, mode, pad, [None if s is None else s.shape for s in srcs]))
        pkt = acb.mixed_frame_packet(list(range(n)), W, H, caps, "standard")
        if exp[0] is not None and pkt[0] != want_hdr(exp[0], W, H) + exp[0]:
            fail("packet", (n, W, H, level, mode, pad))
        for k in range(n):
            acb.source_clear(k)
        counts["mixed+packet"] += 1
    elif fam == 6:
        n = int(rng.integers(1, 12))
        level, mode = int(rng.integers(0, 4)), int(rng.choice([0, 2]))
        cols, rows = int(rng.integers(4, 60)), int(rng.integers(2, 20))
        srcs = [ob.port_convert(ob.gen(PATS[k % 4], 96, 64, k), cols, rows, level, mode) for k in range(n)]
        if rng.random() < 0.2:
            srcs[int(rng.integers(0, n))] = None
        W, H = int(rng.integers(10, 260)), int(rng.integers(3, 80))
        trace("grid", i, n, level, mode, cols, rows, W, H, [None if s is None else len(s) for s in srcs])
        got, exp = acb.ascii_create_grid(srcs, W, H), want_grid(srcs, W, H)
        # When an ANSI spill lands on the canvas terminator the reference returns a string that runs past its own
        # allocation (strlen over the heap: undefined); the library keeps the terminator.  Compare the canvas proper.
        canvas = W * H + H
        if got != exp and not (exp[0] is not None and len(exp[0]) > canvas and got[0] == exp[0][:canvas]):
            fail("grid", (n, level, mode, cols, rows, W, H))
        counts["grid"] += 1
    else:
        img = rnd_img(i)
        f, t = int(rng.integers(-1, 14)), float(rng.random() * 20)
        trace("filter", i, img.shape, f, t)
        a = acb.apply_color_filter(img, f, t)
        b = (ob.ref_color_filter if use_ref else ob.port_color_filter)(img, f, t)
        if a[0] != b[0] or not np.array_equal(a[1], b[1]):
            fail("filter", (img.shape, f, t))
        counts["filter"] += 1
print("fuzz parity ok (%s, seed %d, %.0f s): %s" % ("compiled reference" if use_ref else "port", seed, budget,
                                                     ", ".join("%s %d" % kv for kv in counts.items())))
