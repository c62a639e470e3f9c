"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Checker = the compiled reference (oracle/_ref) when its .so travelled with the snapshot, else the pinned
port.  Bar: byte-for-byte equality of every frame string (the path is all-integer; no tolerance).
"""
import ctypes as C
import itertools
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LEVELS = (0, 1, 2, 3)
MODES = (0, 1, 2)


@pytest.fixture(scope="module")
def acb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ascii_chat_b200 as m
    assert m.lib().acb200_init(0) == 0, m.last_error()
    return m


@pytest.fixture(scope="module")
def want(ob):
    """oracle convert: compiled reference if present, else the port"""
    if ob.ref() is not None:
        return ob.ref_convert
    return ob.port_convert


def _conv(acb, img, c, r, level, mode, pal="standard", aspect=False, stretch=False, pad=False):
    return acb.ascii_convert_with_capabilities(img, c, r, acb.make_caps(level, mode, pad), aspect, stretch, pal)


# ---------------------------------------------------------------- drop-in calls, NN (reference-exact) mode
def test_matrix_vs_oracle(acb, ob, want):
    shapes = [(64, 48, 16, 8), (100, 37, 33, 11), (17, 9, 5, 3), (320, 240, 80, 24), (8, 2, 8, 2), (31, 64, 40, 20)]
    n = 0
    for pat, (W, H, c, r) in itertools.product(("noise", "gradient", "bars", "grey", "solid"), shapes):
        img = ob.gen(pat, W, H, 3)
        for level, mode, pal in itertools.product(LEVELS, MODES, ("standard", "blocks", "digital", "minimal")):
            for aspect, pad in ((False, False), (True, True), (True, False)):
                got = _conv(acb, img, c, r, level, mode, pal, aspect, False, pad)
                exp = want(img, c, r, level, mode, pal, aspect, False, pad)
                assert got == exp, (pat, W, H, c, r, level, mode, pal, aspect, pad)
                n += 1
    assert n > 3000


def test_golden_fingerprints(acb, ob, golden):
    """committed fingerprints of the compiled reference, incl. the BASELINE configs C1/C2/C3 at full size"""
    for rec in golden["frames"]:
        img = ob.gen(rec["pattern"], rec["W"], rec["H"], 0)
        s = _conv(acb, img, rec["cols"], rec["rows"], rec["level"], rec["mode"], rec["palette"], bool(rec["aspect"]),
                  False, bool(rec["pad"]))
        assert s is not None, (rec, acb.last_error())
        assert (len(s), s.count(b"\n"), "%08x" % ob.fnv(s)) == (rec["bytes"], rec["newlines"], rec["fnv"]), rec


def test_low_entropy_images(acb, ob):
    """long runs, black holes, near-grey colours, stripes: the REP / reset / SGR-dedupe rules"""
    rng = np.random.default_rng(7)
    chk = ob.ref_print if ob.ref() is not None else ob.port_print
    for it in range(60):
        w, h = int(rng.integers(1, 70)), int(rng.integers(1, 40))
        kind = it % 4
        if kind == 0:
            img = rng.integers(0, 2, (h, w, 1), dtype=np.uint8).repeat(3, axis=2) * 255
        elif kind == 1:
            img = (rng.integers(0, 3, (h, w, 3)) * 20).astype(np.uint8)
        elif kind == 2:
            img = np.repeat(rng.integers(0, 256, (h, (w + 6) // 7, 3), dtype=np.uint8), 7, axis=1)[:, :w]
        else:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            img[rng.random((h, w)) < 0.4] = 0
        for level, mode, pal in itertools.product(LEVELS, MODES, ("standard", "cool")):
            got = acb.image_print_with_capabilities(img, acb.make_caps(level, mode), pal)
            assert got == chk(img, level, mode, pal), (it, w, h, level, mode, pal)


def test_wide_rows(acb, ob, want):
    """rows wider than one CTA pass (multi-segment scans) and wider than the shared staging buffer"""
    for W, c, r in ((1400, 700, 3), (3840, 1500, 2), (2000, 2000, 2), (3840, 3840, 1)):
        img = ob.gen("noise", W, 16, 1)
        img[:, W // 3: W // 2] = 0
        img[:, : W // 8] = (9, 200, 30)
        for level, mode in itertools.product(LEVELS, (0, 2)):
            assert _conv(acb, img, c, r, level, mode) == want(img, c, r, level, mode), (W, c, r, level, mode)


def test_legacy_ascii_convert(acb, ob):
    img = ob.gen("noise", 120, 90, 1)
    chk = ob.ref_convert_legacy if ob.ref() is not None else ob.port_convert_legacy
    for color, aspect, stretch, opt in itertools.product((False, True), (False, True), (False, True), MODES):
        acb.lib().acb200_set_option_render_mode(opt)
        got = acb.ascii_convert(img, 40, 20, color, aspect, stretch, "standard")
        acb.lib().acb200_set_option_render_mode(0)
        assert got == chk(img, 40, 20, color, aspect, stretch, "standard", opt), (color, aspect, stretch, opt)


def test_image_resize(acb, ob):
    rng = np.random.default_rng(3)
    chk = ob.ref_resize if ob.ref() is not None else ob.port_resize
    for _ in range(25):
        sw, sh, dw, dh = (int(rng.integers(1, 300)) for _ in range(4))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(acb.image_resize(src, dw, dh), chk(src, dw, dh)), (sw, sh, dw, dh)


def test_error_behaviour(acb):
    """NULL / invalid arguments: NULL result + ERROR_INVALID_PARAM, as ascii.c:198-212, 256-265"""
    caps = acb.make_caps(3, 0)
    img = np.zeros((4, 4, 3), np.uint8)
    L = acb.lib()
    assert L.ascii_convert_with_capabilities(None, 4, 4, C.byref(caps), False, False, b"ab") is None
    assert acb.last_error()[0] == 86
    assert acb.ascii_convert_with_capabilities(img, 0, 4, caps, False, False, "standard") is None
    assert acb.last_error()[0] == 86
    assert acb.ascii_convert_with_capabilities(img, 4000, 4, caps, False, False, "standard") is None  # > IMAGE_MAX_WIDTH
    assert acb.ascii_convert_with_capabilities(img, 4, 4, caps, False, False, None) is None
    assert acb.ascii_convert_with_capabilities(img, 4, 4, caps, False, False, "") is None  # empty palette (common.c:275)
    assert acb.ascii_convert(img, 4, 4, True, False, False, "") is None
    # a healthy call still works afterwards
    assert acb.ascii_convert_with_capabilities(img, 4, 4, caps, False, False, "standard") is not None


def test_concurrent_callers(acb, ob, want):
    """one render thread per client (src/server/render.c:340): per-thread streams, shared LUT cache"""
    imgs = [ob.gen("noise", 160, 120, i) for i in range(8)]
    exp = [want(im, 40, 15, 3, 2) for im in imgs]
    errs = []

    def worker(i):
        for _ in range(20):
            if _conv(acb, imgs[i], 40, 15, 3, 2) != exp[i]:
                errs.append(i)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


# ---------------------------------------------------------------- box-filter mode (our spec; oracle = CPU box + oracle print)
def test_box_mode_vs_oracle(acb, ob):
    cases = [(640, 480, 80, 24), (320, 240, 64, 20), (333, 127, 47, 13), (64, 64, 64, 32), (1920, 1080, 160, 48),
             (100, 50, 130, 70), (48, 2000, 7, 5)]  # streaming path, generic path, upscale, tall bands
    for (W, H, c, r), pat in itertools.product(cases, ("noise", "bars")):
        img = ob.gen(pat, W, H, 2)
        for level, mode, pal in ((0, 0, "standard"), (2, 0, "standard"), (3, 0, "standard"), (3, 0, "blocks"),
                                 (3, 0, "cool"), (3, 2, "standard"), (1, 2, "standard"), (0, 2, "standard"),
                                 (2, 2, "standard"), (3, 1, "standard"), (1, 0, "digital")):
            rows_px = r * 2 if mode == 2 else r
            cfg = acb.make_cfg(W, H, c, rows_px, level, mode, pal, scale=acb.SCALE_BOX, pad_left=3, pad_top=2)
            got = acb.render_batch_host(cfg, [img])[0]
            exp = ob.port_convert(img, c, r, level, mode, pal, scale=ob.SCALE_BOX)
            exp = ob._take(ob.port().orc_pad_height(ob._take(ob.port().orc_pad_width(exp, 3)), 2))
            assert got == exp, (W, H, c, r, pat, level, mode, pal)


def test_box_truecolor_fg_colour_carry(acb, ob):
    """the cross-row colour state of ansi_rle_add_pixel through the look-back: flat colours make whole rows drop
    their first SGR, multi-byte glyph rows must be transparent to the carry"""
    rng = np.random.default_rng(2)
    for it in range(12):
        W, H, c, r = 640, 48 * (1 + it % 3), 40 + 8 * (it % 4), 6 + it % 5
        img = np.zeros((H, W, 3), np.uint8)
        band = max(1, H // (r * 2))
        for y0 in range(0, H, band):  # horizontal bands of few colours: rows often start with the colour above
            img[y0:y0 + band] = rng.choice([0, 40, 200, 255], 3)
        img[:, W // 2:] = img[:, W // 2:] // 2 + rng.integers(0, 2, (H, W - W // 2, 1), dtype=np.uint8) * 100
        for pal in ("standard", "blocks", "cool", "minimal"):
            cfg = acb.make_cfg(W, H, c, r, 3, 0, pal, scale=acb.SCALE_BOX)
            got = acb.render_batch_host(cfg, [img, img[::-1].copy()])
            for g, src in zip(got, (img, img[::-1].copy())):
                assert g == ob.port_convert(src, c, r, 3, 0, pal, scale=ob.SCALE_BOX), (it, pal)


# ---------------------------------------------------------------- batch API on resident frames
def _device_batch(acb, frames, cfg):
    import torch
    n = len(frames)
    d_in = torch.from_numpy(np.stack(frames)).cuda()
    cap = acb.frame_capacity(cfg)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    acb.render_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr(), None)
    acb.synchronize()
    tot, ker = acb.time_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(),
                                     d_scr.data_ptr(), 1)
    assert tot > 0 and ker > 0 and ker <= tot * 1.05
    lens = d_len.cpu().numpy()
    out = d_out.cpu().numpy().reshape(n, cap)
    res = []
    for i in range(n):
        assert out[i, lens[i]] == 0  # NUL terminated
        res.append(out[i, : lens[i]].tobytes())
    return res


def test_batch_device_matches_oracle(acb, ob, want):
    frames = [ob.gen(("noise", "bars", "gradient")[i % 3], 320, 240, i) for i in range(7)]
    for level, mode in ((0, 0), (2, 0), (3, 0), (3, 2), (3, 1)):
        cfg = acb.make_cfg(320, 240, 80, 48 if mode == 2 else 24, level, mode)
        res = _device_batch(acb, frames, cfg)
        for i, f in enumerate(frames):
            assert res[i] == want(f, 80, 24, level, mode), (level, mode, i)


def test_full_size_batches_properties(acb, ob):
    """BASELINE shapes at full size: every frame of a resident batch equals its own single-frame render (both
    scalers), strings are NUL-free, have exactly rows-1 newlines, and each text row covers exactly `cols` cells
    once REP sequences are expanded (the consumer-side codec, lib/video/ascii/rle.c)."""
    import re
    for (W, H, c, r, level, mode) in ((1920, 1080, 160, 48, 2, 0), (3840, 2160, 320, 96, 3, 2)):
        frames = [ob.gen("noise" if i % 2 else "bars", W, H, i) for i in range(4)]
        for scale in (acb.SCALE_NN, acb.SCALE_BOX):
            cfg = acb.make_cfg(W, H, c, r * 2 if mode == 2 else r, level, mode, scale=scale)
            res = _device_batch(acb, frames, cfg)
            single = acb.render_batch_host(cfg, frames)
            for i in range(len(frames)):
                assert res[i] == single[i]
                assert b"\0" not in res[i] and res[i].count(b"\n") == r - 1
                exp = ob.port_convert(frames[i], c, r, level, mode, scale=scale)
                assert res[i] == exp
                for line in res[i].split(b"\n"):
                    txt = line.decode("utf-8")
                    txt = re.sub("(.)\x1b\\[(\\d+)b", lambda m: m.group(1) * (int(m.group(2)) + 1), txt)
                    txt = re.sub("\x1b\\[[0-9;]*m", "", txt)
                    assert len(txt) == c


# ---------------------------------------------------------------- grid compositors
def test_text_grid(acb, ob):
    rng = np.random.default_rng(11)
    chk = ob.ref_create_grid if ob.ref() is not None else ob.port_create_grid
    for it in range(40):
        n = int(rng.integers(1, 10))
        level, mode = int(rng.integers(0, 4)), int(rng.choice([0, 2]))
        cols, rows = int(rng.integers(8, 50)), int(rng.integers(3, 16))
        srcs = [ob.port_convert(ob.gen(("noise", "bars", "gradient")[i % 3], 96, 64, i), cols, rows, level, mode)
                for i in range(n)]
        W, H = int(rng.integers(10, 200)), int(rng.integers(3, 60))
        assert acb.ascii_create_grid(srcs, W, H) == chk(srcs, W, H), (it, n, W, H)


def test_text_grid_edge_cases(acb, ob):
    """more than 32 sources (device tables instead of inline ones), missing sources, trailing newlines, more lines
    than the cell is tall, plain text, ANSI spills across cells in narrow grids: every byte as the reference writes it"""
    chk = ob.ref_create_grid if ob.ref() is not None else ob.port_create_grid
    col = [ob.port_convert(ob.gen(("noise", "bars", "gradient")[i % 3], 96, 64, i), 30, 20, 3, 0) for i in range(40)]
    cases = [(col, 300, 80), (col[:33], 200, 70), (col[:5] + [None] + col[6:9], 120, 40),
             ([c + b"\n" for c in col[:4]], 100, 30), ([c + b"\n\n\n" for c in col[:3]], 64, 9),
             ([b"plain\ntext\nonly", b"second\nsource", b"x" * 500], 50, 12), ([b"", col[1]], 80, 24),
             (col[:9], 33, 11), (col[:4], 21, 7), ([col[0]], 12, 4), ([col[0]], 200, 60), ([b"\n\n\n"], 20, 5)]
    for i, (srcs, W, H) in enumerate(cases):
        assert acb.ascii_create_grid(srcs, W, H) == chk(srcs, W, H), (i, len(srcs), W, H)


def test_text_grid_golden(acb, ob, golden):
    for rec in golden["text_grids"]:
        srcs = [ob.port_convert(ob.gen("noise" if i % 2 else "bars", 160, 120, i), rec["cols"], rec["rows"],
                                rec["level"], rec["mode"]) for i in range(rec["n"])]
        g, sz = acb.ascii_create_grid(srcs, rec["W"], rec["H"])
        assert (sz, "%08x" % ob.fnv(g)) == (rec["size"], rec["fnv"]), rec


def test_pixel_composite(acb, ob):
    """server grid: create_multi_source_composite (stream.c:664-779) then the viewer's convert (stream.c:841)"""
    rng = np.random.default_rng(5)
    for it in range(12):
        n = int(rng.integers(1, 10))
        srcs = [ob.gen(("noise", "bars", "gradient")[i % 3], int(rng.integers(40, 400)), int(rng.integers(30, 300)), i)
                for i in range(n)]
        W, H = int(rng.integers(40, 200)), int(rng.integers(20, 60))
        got, gc, gr = acb.composite(srcs, W, H)
        exp, ec, er = (ob.ref_composite if ob.ref() is not None else ob.port_composite)(srcs, W, H)
        assert (gc, gr) == (ec, er) and np.array_equal(got, exp), (it, n, W, H)
        caps = acb.make_caps(3, 2, True)
        a = acb.ascii_convert_with_capabilities(got, W, H * 2, caps, True, False, "standard")
        assert a == ob.port_convert(exp, W, H * 2, 3, 2, "standard", True, False, True)


# ---------------------------------------------------------------- server per-client entry with resident sources
def _load_slots(acb, srcs):
    slots = list(range(len(srcs)))
    for i, s in enumerate(srcs):
        if s is None:
            assert acb.source_clear(i) == 0
        else:
            assert acb.source_update(i, s) == 0, acb.last_error()
    return slots


def test_mixed_frame_golden(acb, ob, golden):
    """acb200_mixed_frame against the reference's create_mixed_ascii_frame_for_client (stream.c:958-1191):
    committed fingerprints of the compiled reference + the port on the same inputs, byte for byte"""
    for rec, case in zip(golden["mixed_frames"], ob.mixed_cases()):
        srcs = ob.mixed_sources(case)
        slots = _load_slots(acb, srcs)
        caps = acb.make_caps(case["level"], case["mode"], bool(case["pad"]))
        s, sz, cnt = acb.mixed_frame(slots, case["W"], case["H"], caps, case["palette"])
        assert (sz, cnt, None if s is None else "%08x" % ob.fnv(s)) == (rec["size"], rec["sources"], rec["fnv"]), case
        exp = ob.port_mixed_frame(srcs, case["W"], case["H"], case["level"], case["mode"], case["palette"],
                                  bool(case["pad"]))
        assert (s, sz, cnt) == exp
    for i in range(acb.MAX_SOURCES):
        acb.source_clear(i)


def test_source_update_wire(acb, ob):
    """frames ingested in their wire form ([w:be32][h:be32][RGB24], protocol.c:737-889) render like the same frames
    ingested as (rgb, w, h); malformed payloads leave the slot untouched"""
    import struct
    srcs = [ob.gen(("noise", "bars", "gradient")[i], 200 + 17 * i, 120 + 9 * i, i) for i in range(3)]
    for i, s_ in enumerate(srcs):
        assert acb.source_update_wire(i, struct.pack(">II", s_.shape[1], s_.shape[0]) + s_.tobytes()) == 0, acb.last_error()
    caps = acb.make_caps(3, 2, True)
    got = acb.mixed_frame([0, 1, 2], 120, 40, caps, "standard")
    assert got == ob.port_mixed_frame(srcs, 120, 40, 3, 2, "standard", True)
    assert acb.source_update_wire(1, struct.pack(">II", 10, 10) + b"short") == 86
    assert acb.mixed_frame([0, 1, 2], 120, 40, caps, "standard") == got  # slot 1 kept its frame
    for i in range(3):
        acb.source_clear(i)


def test_mixed_frame_errors_and_slot_rules(acb, ob):
    caps = acb.make_caps(3, 0, True)
    img = ob.gen("noise", 64, 48, 0)
    for i in range(acb.MAX_SOURCES):
        acb.source_clear(i)
    acb.last_error()
    assert acb.mixed_frame([0, 1], 80, 24, caps, "standard") == (None, 0, 0)  # nobody sends video: no frame, no error
    assert acb.last_error()[0] == 0
    assert acb.mixed_frame([0], 0, 24, caps, "standard")[0] is None and acb.last_error()[0] == 86
    assert acb.mixed_frame([99], 80, 24, caps, "standard")[0] is None and acb.last_error()[0] == 86
    assert acb.source_update(acb.MAX_SOURCES, img) == 86
    acb.last_error()
    # the dimensions collect_video_sources rejects (stream.c:342) drop the client's video
    assert acb.source_update(1, img) == 0
    assert acb.source_update(1, np.zeros((2161, 8, 3), np.uint8)) == 86
    acb.last_error()
    assert acb.mixed_frame([0, 1], 80, 24, caps, "standard") == (None, 0, 0)
    # caps not received yet (stream.c:816): error, no frame
    assert acb.source_update(1, img) == 0
    assert acb.mixed_frame([0, 1], 80, 24, None, "standard")[0] is None and acb.last_error()[0] == 85
    # a source that changes size between frames
    for (w, h) in ((64, 48), (320, 200), (33, 17), (1280, 720)):
        im = ob.gen("bars", w, h, 1)
        assert acb.source_update(1, im) == 0
        assert acb.mixed_frame([0, 1], 80, 24, caps, "standard") == ob.port_mixed_frame([None, im], 80, 24, 3, 0, "standard", True)
    acb.source_clear(1)


def test_mixed_frame_concurrent_render_and_update(acb, ob):
    """one render thread per receiving client (src/server/render.c) while a receive thread keeps replacing one
    sender's frame: every output equals the reference's answer for one of the two frames, never a torn mix"""
    base = [ob.gen(("noise", "bars", "gradient")[i % 3], 320, 240, i) for i in range(4)]
    alt = ob.gen("grey", 200, 150, 9)
    for i, s in enumerate(base):
        assert acb.source_update(i, s) == 0
    views = [(120, 40, 3, 2), (80, 24, 2, 0), (100, 30, 3, 1), (64, 20, 0, 0), (150, 50, 1, 0), (90, 33, 3, 0)]
    exp = []
    for (W, H, level, mode) in views:
        exp.append({ob.port_mixed_frame(v, W, H, level, mode, "standard", True)[0]
                    for v in (base, [base[0], alt] + base[2:])})
    errs, stop = [], threading.Event()

    def receiver():
        k = 0
        while not stop.is_set():
            acb.source_update(1, alt if k & 1 == 0 else base[1])
            k += 1

    def render(j):
        W, H, level, mode = views[j]
        caps = acb.make_caps(level, mode, True)
        for _ in range(40):
            s, sz, cnt = acb.mixed_frame([0, 1, 2, 3], W, H, caps, "standard")
            if s not in exp[j] or cnt != 4:
                errs.append(j)

    rt = threading.Thread(target=receiver)
    rt.start()
    ts = [threading.Thread(target=render, args=(j,)) for j in range(len(views))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    stop.set()
    rt.join()
    for i in range(4):
        acb.source_clear(i)
    assert not errs


# ---------------------------------------------------------------- client display path (display.c:484-671)
def _want_display(ob):
    return ob.ref_display_convert if ob.ref() is not None else ob.port_display_convert


def _display(acb, img, cols, rows, level, mode, palette="standard", aspect=False, stretch=False, pad=False,
             flip_x=False, flip_y=False, color_filter=0, time_s=0.0):
    return acb.display_convert(img, cols, rows, acb.make_caps(level, mode, pad), aspect, stretch, palette, flip_x,
                               flip_y, color_filter, time_s)


def test_display_convert_golden(acb, ob, golden):
    """flip + colour filter + convert + rainbow replace in one pass against fingerprints of the reference's own
    functions run in display.c's order, incl. the 4K / 1080p BASELINE shapes"""
    want = _want_display(ob)
    for rec, case in zip(golden["display_frames"], ob.display_cases()):
        img = ob.gen(case["pattern"], case["W"], case["H"], 0)
        s = _display(acb, img, **ob.display_args(case))
        got = (None, None) if s is None else (len(s), "%08x" % ob.fnv(s))
        assert got == (rec["bytes"], rec["fnv"]), case
        if case["W"] * case["H"] <= 400 * 300:
            assert s == want(img, **ob.display_args(case)), case


def test_display_convert_matrix(acb, ob):
    want = _want_display(ob)
    n = 0
    for pat, (W, H, c, r) in itertools.product(("noise", "bars", "grey"), ((64, 48, 16, 8), (101, 37, 33, 11), (2, 2, 5, 3))):
        img = ob.gen(pat, W, H, 5)
        for level, mode in itertools.product(LEVELS, MODES):
            for filt, (fx, fy) in itertools.product((0, 1, 3, 9, 11, 12, 13), ((0, 0), (1, 0), (0, 1), (1, 1))):
                kw = dict(cols=c, rows=r, level=level, mode=mode, palette="standard" if n % 2 else "blocks",
                          aspect=bool(n % 3 == 0), pad=bool(n % 3 == 0), flip_x=bool(fx), flip_y=bool(fy),
                          color_filter=filt, time_s=0.31 * (n % 17))
                assert _display(acb, img, **kw) == want(img, **kw), kw
                n += 1
    assert n > 3000
    # rainbow colour and the hue function itself (host float, like the reference)
    for t in (0.0, 0.6, 1.75, 2.9, 3.5, 1234.567):
        assert acb.calculate_rainbow(t) == ob.rainbow_rgb(ob.port().orc_calculate_rainbow, t)
    assert acb.calculate_rainbow(0.0) == (255, 21, 21)  # pure red lifted to luminance 120 (color_filter.c:221-235)


def test_display_ops_in_box_mode(acb, ob):
    """our box-filter spec over the flipped, filtered image (generic kernel when a filter is set, the streaming
    kernels with mirrored bands / column ranges otherwise), against the port"""
    for (W, H, c, r) in ((640, 480, 80, 24), (333, 127, 47, 13), (1920, 1080, 160, 48)):
        img = ob.gen("noise", W, H, 2)
        for level, mode in ((3, 2), (2, 0), (3, 0), (0, 0)):
            for filt, fx, fy in ((0, 1, 0), (0, 0, 1), (0, 1, 1), (3, 0, 0), (1, 1, 1), (12, 1, 0)):
                cfg = acb.make_cfg(W, H, c, r * 2 if mode == 2 else r, level, mode, scale=acb.SCALE_BOX, flip_x=fx,
                                   flip_y=fy, color_filter=filt, filter_time=2.2)
                got = acb.render_batch_host(cfg, [img])[0]
                exp = ob.port_display_convert(img, c, r, level, mode, flip_x=fx, flip_y=fy, color_filter=filt,
                                              time_s=2.2, scale=ob.SCALE_BOX)
                assert got == exp, (W, H, level, mode, filt, fx, fy)
                res = _device_batch(acb, [img, img], cfg)
                assert res[0] == exp and res[1] == exp


def test_color_filter_whole_image(acb, ob, golden):
    """apply_color_filter (color_filter.c:274-346) as a device map: golden fingerprints, every filter, sizes whose
    byte count is not a multiple of 48, strided rows, error returns"""
    want = ob.ref_color_filter if ob.ref() is not None else ob.port_color_filter
    img = ob.gen("noise", 333, 127, 0)
    for rec in golden["color_filter"]:
        rc, out = acb.apply_color_filter(img, rec["filter"], rec["time"])
        assert (rc, "%08x" % ob.fnv(out.tobytes())) == (rec["rc"], rec["fnv"]), rec
    for it, (W, H) in enumerate(((1, 1), (5, 3), (16, 1), (17, 2), (640, 480), (1920, 1080))):
        src = ob.gen("noise" if it % 2 else "gradient", W, H, it)
        for f in (-1, 0, 1, 2, 7, 12, 13):
            a, b = acb.apply_color_filter(src, f, 0.9 * it), want(src, f, 0.9 * it)
            assert a[0] == b[0] and (a[1] == b[1]).all(), (W, H, f)
    # strided rows: only the pixel bytes of each row change
    wide = ob.gen("noise", 40, 9, 1)
    view = np.ascontiguousarray(wide[:, :31, :])
    buf = wide.copy()
    assert acb.lib().apply_color_filter(buf.ctypes.data, 31, 9, 40 * 3, 5, 0.0) == 0
    assert (buf[:, :31, :] == want(view, 5, 0.0)[1]).all() and (buf[:, 31:, :] == wide[:, 31:, :]).all()
    assert acb.lib().apply_color_filter(None, 4, 4, 12, 3, 0.0) == -1
    assert acb.lib().apply_color_filter(buf.ctypes.data, 0, 4, 12, 3, 0.0) == -1
    # device-resident image
    import torch
    big = ob.gen("noise", 3840, 2160, 4)
    d = torch.from_numpy(big).cuda()
    torch.cuda.synchronize()
    assert acb.color_filter_device(d.data_ptr(), 3840, 2160, 3840 * 3, 3) == 0
    acb.synchronize()
    assert (d.cpu().numpy() == ob.port_color_filter(big, 3)[1]).all()


def test_reference_unit_test_kats_on_the_device(acb, ob):
    """the known answers the reference's own unit tests hold for the 8f rows (tests/unit/video/color_filter_test.c,
    tests/unit/network/crc32_hw_test.c), asked of the device path"""
    import test_oracle
    import torch
    test_oracle.check_color_filter_kats(acb.apply_color_filter)
    for w, h, st in ((0, 1, 3), (1, 0, 3), (1, 1, 0)):
        buf = np.full(3, 255, np.uint8)
        assert acb.lib().apply_color_filter(buf.ctypes.data, w, h, st, 3, 0.0) == -1
    msgs = [b"", b"Hello, World!", b"\x42", b"ascii-chat", bytes(range(256)) * 4, b"\0" * 100, b"\xff" * 100]
    pitch = 1040
    arena = np.zeros((len(msgs), pitch), np.uint8)
    for i, m in enumerate(msgs):
        arena[i, :len(m)] = np.frombuffer(m, np.uint8)
    d_out = torch.from_numpy(arena).cuda()
    d_len = torch.tensor([len(m) for m in msgs], dtype=torch.int32, device="cuda")
    d_hdr = torch.zeros(len(msgs) * 24, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), len(msgs), 80, 24, d_hdr.data_ptr(), None)
    acb.synchronize()
    hdr = d_hdr.cpu().numpy().reshape(len(msgs), 24)
    crcs = [int.from_bytes(hdr[i, 16:20].tobytes(), "big") for i in range(len(msgs))]
    assert crcs[0] == 0 and crcs[1] == 0x4D551068 and crcs[2] != 0          # crc32_hw_test.c:14-50
    for i, m in enumerate(msgs):
        assert crcs[i] == ob.port().orc_crc32c(m, len(m))


# ---------------------------------------------------------------- wire packaging (acip/server.c:188-236, crc32.c)
def test_frame_packets_device(acb, ob):
    """CRC32-C + ascii_frame_packet_t headers of a resident batch against the oracle's header for the same strings"""
    import torch
    want = ob.ref_packet_header if ob.ref() is not None else ob.port_packet_header
    for (W, H, c, r, level, mode, n) in ((320, 240, 80, 24, 3, 2, 6), (640, 480, 80, 24, 0, 0, 3),
                                         (1920, 1080, 160, 48, 2, 0, 5), (3840, 2160, 320, 96, 3, 2, 3)):
        frames = [ob.gen(("noise", "bars", "gradient")[i % 3], W, H, i) for i in range(n)]
        cfg = acb.make_cfg(W, H, c, r * 2 if mode == 2 else r, level, mode)
        cap = acb.frame_capacity(cfg)
        d_in = torch.from_numpy(np.stack(frames)).cuda()
        d_out = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
        d_hdr = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        acb.render_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr(), None)
        acb.frame_packets_device(d_out.data_ptr(), cap, d_len.data_ptr(), n, c, r, d_hdr.data_ptr(), None)
        acb.synchronize()
        lens, out, hdr = d_len.cpu().numpy(), d_out.cpu().numpy().reshape(n, cap), d_hdr.cpu().numpy().reshape(n, 24)
        for i in range(n):
            assert hdr[i].tobytes() == want(out[i, :lens[i]].tobytes(), c, r), (W, H, level, mode, i)


@pytest.mark.parametrize("crc_form", [1, 2, 0])
def test_crc32c_lengths_and_fixup_device(acb, ob, golden, crc_form):
    """arbitrary byte strings in a device arena: chunk-boundary lengths (64 KB chunks, 256-byte segments, 512-byte
    rows), empty frames, the golden CRCs — through the row form (1), the segment form (2) and the size-based choice
    (0); and the trailing-reset cut of stream.c:1085-1127 on synthetic strings"""
    import torch
    acb.lib().acb200_set_crc_form(crc_form)
    lens = [0, 1, 3, 63, 64, 65, 127, 128, 255, 256, 257, 319, 320, 4095, 16383, 16384, 16385, 32768, 65535, 65536,
            65537, 65600, 100001, 131072, 1180548,
            15, 16, 17, 511, 512, 513, 527, 528, 1023, 1024, 1025, 2047, 2048, 2049, 4096 + 16, 512 * 5 - 1, 512 * 5 + 31]
    pitch = (max(lens) + 1 + 15) & ~15
    arena = np.zeros((len(lens), pitch), np.uint8)
    data = []
    for i, L in enumerate(lens):
        d = ob.gen("noise", max(1, (L + 2) // 3), 1, 7).tobytes()[:L]
        arena[i, :L] = np.frombuffer(d, np.uint8)
        data.append(d)
    d_out = torch.from_numpy(arena).cuda()
    d_len = torch.tensor(lens, dtype=torch.int32, device="cuda")
    d_hdr = torch.zeros(len(lens) * 24, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), len(lens), 320, 96, d_hdr.data_ptr(), None)
    acb.synchronize()
    hdr = d_hdr.cpu().numpy().reshape(len(lens), 24)
    by_len = {rec["len"]: rec for rec in golden["crc32c"] if "len" in rec}
    for i, L in enumerate(lens):
        assert hdr[i].tobytes() == ob.port_packet_header(data[i], 320, 96), L
        if L in by_len:
            assert hdr[i].tobytes().hex() == by_len[L]["header"], L

    # many frames of unrelated lengths in one arena: the row kernel's warps take equal shares of ROWS, so their runs
    # start and end anywhere, straddle frames and skip frames without a full row; the plan's scan spans several blocks
    rng = np.random.default_rng(5)
    for n, hi in ((2500, 3000), (37, 70000), (1, 5000), (300, 700)):
        lens2 = rng.integers(0, hi, n)
        pitch = (int(lens2.max()) + 1 + 15) & ~15
        arena = rng.integers(0, 256, (n, pitch), dtype=np.uint8)
        d_out = torch.from_numpy(arena).cuda()
        d_len = torch.tensor(lens2, dtype=torch.int32, device="cuda")
        d_hdr = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), n, 7, 9, d_hdr.data_ptr(), None)
        acb.synchronize()
        hdr = d_hdr.cpu().numpy().reshape(n, 24)
        for i in range(n):
            m = arena[i, :lens2[i]].tobytes()
            assert int.from_bytes(hdr[i, 16:20].tobytes(), "big") == ob.port().orc_crc32c(m, len(m)), (n, i, lens2[i])
            assert int.from_bytes(hdr[i, 8:12].tobytes(), "big") == lens2[i]

    acb.lib().acb200_set_crc_form(0)
    rst = b"\x1b[0m"
    cases = [b"", b"abc", rst, b"row" + rst, b"ab" + rst + b"cd", rst + b"x" * 1000, b"y" * 700 + rst + b"z" * 3,
             b"plain text without any reset" * 40, (b"q" + rst) * 300 + b"tail", b"\x1b[0", b"a" + rst + rst + b"\x1b[0"]
    pitch = 4096
    arena = np.zeros((len(cases), pitch), np.uint8)
    for i, s in enumerate(cases):
        arena[i, :len(s)] = np.frombuffer(s, np.uint8)
    d_out = torch.from_numpy(arena).cuda()
    d_len = torch.tensor([len(s) for s in cases], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    acb.trailing_reset_fixup_device(d_out.data_ptr(), pitch, d_len.data_ptr(), len(cases), None)
    acb.synchronize()
    got_len, got = d_len.cpu().numpy(), d_out.cpu().numpy()
    for i, s in enumerate(cases):
        exp = ob.mixed_frame_fixup(s)
        assert got[i, :got_len[i]].tobytes() == exp, s[:40]


def test_mixed_frame_packet(acb, ob, golden):
    """create_mixed_ascii_frame_for_client + acip_send_ascii_frame's header in one call: header||frame equals the
    oracle's header for the reference's frame, for every golden server case"""
    for rec, case in zip(golden["mixed_frames"], ob.mixed_cases()):
        srcs = ob.mixed_sources(case)
        slots = _load_slots(acb, srcs)
        caps = acb.make_caps(case["level"], case["mode"], bool(case["pad"]))
        pkt, sz, cnt = acb.mixed_frame_packet(slots, case["W"], case["H"], caps, case["palette"])
        frame, fsz, fcnt = acb.mixed_frame(slots, case["W"], case["H"], caps, case["palette"])
        assert cnt == fcnt == rec["sources"]
        if frame is None:
            assert pkt is None and sz == 0
            continue
        assert "%08x" % ob.fnv(frame) == rec["fnv"]
        assert sz == 24 + fsz and pkt[24:] == frame, case
        assert pkt[:24] == ob.port_packet_header(frame, case["W"], case["H"]), case
    for i in range(acb.MAX_SOURCES):
        acb.source_clear(i)


def test_concurrent_callers_new_entries(acb, ob):
    """the §8f entry points from many threads at once (each thread leases its own stream / staging): display
    conversions with different flips and filters, packet-producing server frames, grids — all byte-exact"""
    img = ob.gen("noise", 320, 240, 1)
    srcs = [ob.gen(("noise", "bars", "gradient")[i % 3], 160, 120, i) for i in range(3)]
    for i, s in enumerate(srcs):
        assert acb.source_update(i, s) == 0
    grid_srcs = [ob.port_convert(ob.gen("bars", 96, 64, i), 20, 8, 3, 0) for i in range(4)]
    jobs = []
    for j in range(12):
        kw = dict(cols=40 + j, rows=12 + j % 5, level=(3, 2, 1, 0)[j % 4], mode=(2, 0, 1)[j % 3], flip_x=bool(j & 1),
                  flip_y=bool(j & 2), color_filter=(0, 3, 12, 1)[j % 4], time_s=0.5 * j)
        jobs.append(("display", kw, ob.port_display_convert(img, **kw)))
    for j in range(6):
        W, H, level, mode = 60 + 7 * j, 20 + j, (3, 2, 0)[j % 3], (2, 0)[j % 2]
        frame = ob.port_mixed_frame(srcs, W, H, level, mode, "standard", True)[0]
        jobs.append(("packet", (W, H, level, mode), ob.port_packet_header(frame, W, H) + frame))
    for j in range(4):
        jobs.append(("grid", (50 + 10 * j, 20 + j), ob.port_create_grid(grid_srcs, 50 + 10 * j, 20 + j)))
    errs = []

    def run(kind, arg, exp):
        for _ in range(15):
            if kind == "display":
                got = _display(acb, img, **arg)
            elif kind == "packet":
                W, H, level, mode = arg
                got = acb.mixed_frame_packet([0, 1, 2], W, H, acb.make_caps(level, mode, True), "standard")[0]
            else:
                got = acb.ascii_create_grid(grid_srcs, *arg)
            if got != exp:
                errs.append((kind, arg))
                return

    ts = [threading.Thread(target=run, args=j) for j in jobs]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(3):
        acb.source_clear(i)
    assert not errs, errs[:3]


def test_display_ops_on_wide_rows(acb, ob):
    """the rows-too-wide-for-shared-staging path (scratch rows + k_stitch) with the display steps: the rainbow colour's
    digit count drives the conditional first-SGR drop of truecolor-fg across rows; flips and filters ride along"""
    want = _want_display(ob)
    for W, c, r in ((3840, 3000, 3), (2600, 2600, 2)):
        img = ob.gen("noise", W, 12, 2)
        img[:, : W // 6] = (40, 40, 40)        # long equal-colour stretches: SGR dedupe across rows matters
        img[:, W // 2: W // 2 + 300] = 0
        for level, mode in ((3, 0), (3, 2), (2, 0)):
            for filt, fx, fy, t in ((12, 1, 0, 0.9), (12, 0, 1, 2.6), (3, 1, 1, 0.0), (0, 1, 0, 0.0)):
                kw = dict(cols=c, rows=r, level=level, mode=mode, flip_x=bool(fx), flip_y=bool(fy), color_filter=filt,
                          time_s=t)
                assert _display(acb, img, **kw) == want(img, **kw), (W, c, r, level, mode, filt, fx, fy)
