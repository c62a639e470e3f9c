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
counts = {"display": 0, "mixed+packet": 0, "grid": 0, "filter": 0}
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
    fam = i % 8
    if fam < 5:
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
