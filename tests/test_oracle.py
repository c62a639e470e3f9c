"""CPU tests that PIN the oracle (oracle/ascii_oracle.c, "the port").

Three anchors, strongest first:
  1. the compiled reference itself (oracle/_ref, built from /root/reference by oracle/Makefile)
     on randomised matrices — skipped only where that library was never built;
  2. committed fingerprints of the compiled reference's output (tests/golden/reference_vectors.json,
     made by tests/golden/make_golden.py), incl. the survey's Appendix-C anchors;
  3. the known-answer values the reference's own unit tests hold for this path
     (tests/unit/util/ansi_fast_test.c, output_buffer_test.c, aspect_ratio_test.c).
"""
import ctypes as C
import itertools

import numpy as np
import pytest

LEVELS = (0, 1, 2, 3)
MODES = (0, 1, 2)


# ----------------------------------------------------------------- 2. golden fingerprints
def test_port_matches_golden_frames(ob, golden):
    bad = []
    for rec in golden["frames"]:
        img = ob.gen(rec["pattern"], rec["W"], rec["H"], 0)
        s = ob.port_convert(img, rec["cols"], rec["rows"], rec["level"], rec["mode"], rec["palette"],
                            bool(rec["aspect"]), False, bool(rec["pad"]))
        got = (len(s), s.count(b"\n"), "%08x" % ob.fnv(s))
        if got != (rec["bytes"], rec["newlines"], rec["fnv"]):
            bad.append((rec, got))
    assert not bad, bad[:3]


def test_port_quantiser_tables_match_golden(ob, golden):
    tab = np.empty(1 << 24, np.uint8)
    p = tab.ctypes.data_as(C.POINTER(C.c_uint8))
    ob.port().orc_fill_table(0, None, p)
    assert "%08x" % ob.fnv(tab.tobytes()) == golden["rgb_to_256color_table_fnv"]
    ob.port().orc_fill_table(1, None, p)
    assert "%08x" % ob.fnv(tab.tobytes()) == golden["rgb_to_16color_table_fnv"]


def test_port_quirk_literals(ob, golden):
    """SURVEY.md §8a Q1-Q5 on an all-white 8x2 image."""
    white = np.full((2, 8, 3), 255, np.uint8)
    q = golden["quirks"]
    assert ob.port_print(white, 0, 0) == q["Q1_mono"].encode("latin-1") == b";\x1b[7b\n;\x1b[7b"
    assert ob.port_print(white, 1, 0) == q["Q2_16"].encode("latin-1")
    assert ob.port_print(white, 2, 0) == q["Q3_256"].encode("latin-1")
    assert ob.port_print(white, 3, 0) == q["Q3_true"].encode("latin-1")
    assert ob.port_print(white, 3, 1) == q["Q4_true_bg"].encode("latin-1")
    assert ob.port_print(white, 3, 2) == q["Q5_half"].encode("latin-1")


def test_port_glyph_tables_match_golden(ob, golden):
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)
    for pal, h in golden["glyph_tables"].items():
        assert "%08x" % ob.fnv(ob.port_print(ramp, 0, 0, pal)) == h["mono"], pal
        assert "%08x" % ob.fnv(ob.port_print(ramp, 1, 0, pal)) == h["c16"], pal
        assert "%08x" % ob.fnv(ob.port_print(ramp, 2, 0, pal)) == h["c256"], pal
        assert "%08x" % ob.fnv(ob.port_print(ramp, 3, 0, pal)) == h["true"], pal


def _nn_src(sw, sh):
    src = np.zeros((sh, sw, 3), np.uint8)
    xs = np.arange(sw, dtype=np.uint32)
    ys = np.arange(sh, dtype=np.uint32)
    src[:, :, 0] = (xs & 255)[None, :]
    src[:, :, 1] = ((xs >> 8) & 255)[None, :] | ((ys[:, None] >> 8) << 6).astype(np.uint8)
    src[:, :, 2] = (ys & 255)[:, None]
    return src


def test_port_nn_resize_matches_golden(ob, golden):
    for rec in golden["nn_resize"]:
        out = ob.port_resize(_nn_src(rec["sw"], rec["sh"]), rec["dw"], rec["dh"])
        assert "%08x" % ob.fnv(out.tobytes()) == rec["fnv"], rec


def test_port_text_grid_matches_golden(ob, golden):
    for rec in golden["text_grids"]:
        srcs = [ob.port_convert(ob.gen("noise" if i % 2 else "bars", 160, 120, i), rec["cols"], rec["rows"],
                                rec["level"], rec["mode"]) for i in range(rec["n"])]
        g, sz = ob.port_create_grid(srcs, rec["W"], rec["H"])
        assert (sz, "%08x" % ob.fnv(g)) == (rec["size"], rec["fnv"]), rec


# ----------------------------------------------------------------- 3. the reference's own KATs
def test_kat_sgr_strings(ob):
    """ansi_fast_test.c:60,102,300,312,316,383-455 — observed through 1x1 renders of the port."""
    px = lambda r, g, b: np.array([[[r, g, b]]], np.uint8)  # noqa: E731
    s = ob.port_print(px(255, 128, 64), 3, 0)
    assert s.startswith(b"\033[38;2;255;128;64m") and s.endswith(b"\033[0m")
    s = ob.port_print(np.array([[[255, 128, 64]], [[100, 200, 50]]], np.uint8), 3, 2)
    assert s == b"\033[38;2;255;128;64m\033[48;2;100;200;50m\xe2\x96\x80\033[0m"
    assert ob.port_print(px(255, 255, 255), 2, 0).startswith(b"\033[38;5;255m")
    assert ob.port_print(px(0, 0, 0), 2, 0).startswith(b"\033[38;5;232m")
    assert ob.port_print(px(128, 0, 0), 1, 0).startswith(b"\033[31m")
    assert ob.port_print(px(0, 128, 0), 1, 0).startswith(b"\033[32m")
    assert ob.port_print(px(255, 255, 255), 1, 0).startswith(b"\033[97m")


def test_kat_rgb_to_16color(ob):
    """ansi_fast_test.c:458-492"""
    P = ob.port()
    for rgb, idx in (((255, 0, 0), 9), ((0, 255, 0), 10), ((0, 0, 255), 12), ((0, 0, 0), 0), ((255, 255, 255), 15),
                     ((128, 0, 0), 1), ((0, 128, 0), 2), ((0, 0, 128), 4), ((192, 192, 192), 7)):
        assert P.orc_rgb_to_16(*rgb) == idx


def test_kat_rgb_to_256color_ranges(ob):
    """ansi_fast_test.c:319-366 only range-checks: greys -> 232..255, colours -> cube 16..231"""
    P = ob.port()
    for v in range(256):
        assert 232 <= P.orc_rgb_to_256(v, v, v) <= 255
    for rgb in ((255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (40, 200, 90)):
        assert 16 <= P.orc_rgb_to_256(*rgb) <= 231


def test_kat_rep_and_digits(ob):
    """output_buffer_test.c:297-305, 337-350"""
    P = ob.port()
    assert [bool(P.orc_rep_is_profitable(n)) for n in (0, 1, 2, 3, 4, 5, 6, 10, 100)] == [False] * 6 + [True] * 3
    for v, d in ((0, 1), (9, 1), (10, 2), (99, 2), (100, 3), (999, 3), (1000, 4), (9999, 4), (10000, 5),
                 (100000, 6), (1000000, 7), (10000000, 8), (100000000, 9), (1000000000, 10), (4294967295, 10)):
        assert P.orc_digits_u32(v) == d


def _aspect(ob, iw, ih, w, h, stretch=False):
    ow, oh = C.c_long(0), C.c_long(0)
    ob.port().orc_aspect_ratio(iw, ih, w, h, int(stretch), C.byref(ow), C.byref(oh))
    return ow.value, oh.value


def test_kat_aspect_ratio(ob):
    """aspect_ratio_test.c:36-39 (stretch), :100-103 (degenerate -> 1x1); SURVEY §8a a2 worked values"""
    assert _aspect(ob, 1920, 1080, 80, 24, True) == (80, 24)
    for iw, ih in ((0, 1080), (1920, 0), (-1920, 1080), (1920, -1080)):
        assert _aspect(ob, iw, ih, 80, 24) == (1, 1)
    assert _aspect(ob, 3840, 2160, 320, 96) == (320, 90)
    assert _aspect(ob, 1920, 1080, 160, 48) == (160, 45)


# ----------------------------------------------------------------- 1. against the compiled reference
def test_port_vs_ref_matrix(ob, ref_lib):
    shapes = [(64, 48, 16, 8), (100, 37, 33, 11), (17, 9, 5, 3), (320, 240, 80, 24), (8, 2, 8, 2), (31, 64, 40, 20)]
    n = 0
    for pat, (W, H, c, r) in itertools.product(("noise", "gradient", "bars", "grey", "solid"), shapes):
        img = ob.gen(pat, W, H, 3)
        for level, mode, pal in itertools.product(LEVELS, MODES, ("standard", "blocks", "digital", "minimal")):
            for aspect, pad in ((False, False), (True, True), (True, False)):
                a = ob.ref_convert(img, c, r, level, mode, pal, aspect, False, pad)
                b = ob.port_convert(img, c, r, level, mode, pal, aspect, False, pad)
                assert a == b, (pat, W, H, c, r, level, mode, pal, aspect, pad)
                n += 1
    assert n > 3000


def test_port_vs_ref_random_images(ob, ref_lib):
    """low-entropy random images: long runs, black holes, near-grey colours, 2-colour stripes"""
    rng = np.random.default_rng(7)
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
        for level, mode in itertools.product(LEVELS, MODES):
            for pal in ("standard", "cool"):
                assert ob.ref_print(img, level, mode, pal) == ob.port_print(img, level, mode, pal), (it, level, mode)


def test_port_vs_ref_legacy_ascii_convert(ob, ref_lib):
    img = ob.gen("noise", 120, 90, 1)
    for color, aspect, stretch, opt in itertools.product((False, True), (False, True), (False, True), MODES):
        a = ob.ref_convert_legacy(img, 40, 20, color, aspect, stretch, "standard", opt)
        b = ob.port_convert_legacy(img, 40, 20, color, aspect, stretch, "standard", opt)
        assert a == b, (color, aspect, stretch, opt)


def test_port_vs_ref_resize_and_aspect(ob, ref_lib):
    rng = np.random.default_rng(3)
    for _ in range(40):
        sw, sh, dw, dh = (int(rng.integers(1, 300)) for _ in range(4))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(ob.ref_resize(src, dw, dh), ob.port_resize(src, dw, dh)), (sw, sh, dw, dh)
    for _ in range(3000):
        iw, ih, w, h = (int(rng.integers(1, 4000)) for _ in range(4))
        ow, oh = C.c_ssize_t(0), C.c_ssize_t(0)
        ref_lib.aspect_ratio(iw, ih, w, h, False, C.byref(ow), C.byref(oh))
        assert (ow.value, oh.value) == _aspect(ob, iw, ih, w, h), (iw, ih, w, h)


def test_port_vs_ref_quantiser_tables(ob, ref_lib, golden):
    a = np.empty(1 << 24, np.uint8)
    b = np.empty(1 << 24, np.uint8)
    u8p = C.POINTER(C.c_uint8)
    for which, fn in ((0, ref_lib.rgb_to_256color), (1, ref_lib.rgb_to_16color)):
        ob.port().orc_fill_table(2, C.cast(fn, C.c_void_p), a.ctypes.data_as(u8p))
        ob.port().orc_fill_table(which, None, b.ctypes.data_as(u8p))
        assert np.array_equal(a, b)


def test_port_vs_ref_text_grid(ob, ref_lib):
    rng = np.random.default_rng(11)
    for it in range(40):
        n = int(rng.integers(1, 10))
        level, mode = int(rng.integers(0, 4)), int(rng.choice([0, 2]))
        cols, rows = int(rng.integers(8, 50)), int(rng.integers(3, 16))
        srcs = [ob.ref_convert(ob.gen(("noise", "bars", "gradient")[i % 3], 96, 64, i), cols, rows, level, mode)
                for i in range(n)]
        W, H = int(rng.integers(10, 200)), int(rng.integers(3, 60))
        a = ob.ref_create_grid(srcs, W, H)
        b = ob.port_create_grid(srcs, W, H)
        assert a == b, (it, n, W, H)


def test_port_vs_ref_padding(ob, ref_lib):
    s = b"ab\ncd\n\nxyz"
    for pad in (0, 1, 5):
        a = ob._take(ref_lib.ascii_pad_frame_width(s, pad))
        assert a == ob._take(ob.port().orc_pad_width(s, pad))
        a = ob._take(ref_lib.ascii_pad_frame_height(s, pad))
        assert a == ob._take(ob.port().orc_pad_height(s, pad))


def test_box_filter_spec(ob):
    """our box-filter specification: identity at 1:1, exact means, rounding half up"""
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    assert np.array_equal(ob.port_resize(src, 64, 48, ob.SCALE_BOX), src)
    out = ob.port_resize(src, 16, 12, ob.SCALE_BOX)
    blk = src.reshape(12, 4, 16, 4, 3).astype(np.uint32).sum(axis=(1, 3))
    assert np.array_equal(out, ((blk + 8) // 16).astype(np.uint8))
    # ragged boundaries: floor(dx*sw/dw) partition covers every source pixel exactly once
    src = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    out = ob.port_resize(src, 10, 7, ob.SCALE_BOX)
    xs = [(d * 53) // 10 for d in range(11)]
    ys = [(d * 37) // 7 for d in range(8)]
    for dy in range(7):
        for dx in range(10):
            b = src[ys[dy]:ys[dy + 1], xs[dx]:xs[dx + 1]].astype(np.uint32)
            n = b.shape[0] * b.shape[1]
            assert np.array_equal(out[dy, dx], ((b.sum(axis=(0, 1)) + n // 2) // n).astype(np.uint8))


def _composite_cases():
    rng = np.random.default_rng(5)
    for it in range(30):
        n = int(rng.integers(1, 10))
        srcs = [("noise", "bars", "gradient")[i % 3] for i in range(n)]
        dims = [(int(rng.integers(40, 400)), int(rng.integers(30, 300))) for _ in range(n)]
        yield it, srcs, dims, int(rng.integers(40, 200)), int(rng.integers(20, 60))


def test_port_vs_ref_pixel_composite(ob, ref_lib):
    """server grid: the reference's own create_multi_source_composite (stream.c:664-779), compiled into oracle/_ref"""
    for it, pats, dims, W, H in _composite_cases():
        srcs = [ob.gen(p, w, h, i) for i, (p, (w, h)) in enumerate(zip(pats, dims))]
        a, ac, ar = ob.ref_composite(srcs, W, H)
        b, bc, br = ob.port_composite(srcs, W, H)
        assert (ac, ar) == (bc, br) and np.array_equal(a, b), (it, W, H)


def test_port_pixel_composite_matches_golden(ob, golden):
    for rec, (it, pats, dims, W, H) in zip(golden["pixel_composites"], _composite_cases()):
        srcs = [ob.gen(p, w, h, i) for i, (p, (w, h)) in enumerate(zip(pats, dims))]
        b, bc, br = ob.port_composite(srcs, W, H)
        assert (bc, br, "%08x" % ob.fnv(b.tobytes())) == (rec["cols"], rec["rows"], rec["fnv"]), rec


# ----------------------------------------------------------------- server per-client entry (stream.c:958-1191)
def test_port_mixed_frame_matches_golden(ob, golden):
    """the port's composition of the server entry against fingerprints of the reference's own
    create_mixed_ascii_frame_for_client, run end to end inside oracle/_ref"""
    recs = golden["mixed_frames"]
    cases = ob.mixed_cases()
    assert len(recs) == len(cases)
    for rec, case in zip(recs, cases):
        assert rec["clients"] == case["clients"] and rec["W"] == case["W"]
        s, sz, cnt = ob.port_mixed_frame(ob.mixed_sources(case), case["W"], case["H"], case["level"], case["mode"],
                                         case["palette"], bool(case["pad"]))
        got = (sz, cnt, None if s is None else "%08x" % ob.fnv(s))
        assert got == (rec["size"], rec["sources"], rec["fnv"]), case


def test_port_vs_ref_mixed_frame(ob, ref_lib):
    rng = np.random.default_rng(3)
    done = 0
    while done < 40:
        n = int(rng.integers(1, 11))
        srcs = [None if rng.random() < 0.2 else
                ob.gen(("noise", "bars", "gradient", "grey")[i % 4], int(rng.integers(20, 400)), int(rng.integers(16, 300)), i)
                for i in range(n)]
        W, H = int(rng.integers(20, 200)), int(rng.integers(8, 60))
        level, mode, pad = int(rng.integers(0, 4)), int(rng.integers(0, 3)), bool(rng.integers(0, 2))
        if ob.composite_degenerate(srcs, W, H):
            continue
        done += 1
        assert ob.ref_mixed_frame(srcs, W, H, level, mode, "standard", pad) == \
            ob.port_mixed_frame(srcs, W, H, level, mode, "standard", pad), (n, W, H, level, mode, pad)


def test_mixed_frame_reset_fixup(ob):
    assert ob.mixed_frame_fixup(b"ab\x1b[0mcd") == b"ab\x1b[0m"
    assert ob.mixed_frame_fixup(b"ab\x1b[0m") == b"ab\x1b[0m"
    assert ob.mixed_frame_fixup(b"plain text") == b"plain text"
    assert ob.mixed_frame_fixup(b"x") == b"x"


# ----------------------------------------------------------------- client display path (display.c:484-671)
def test_port_display_matches_golden(ob, golden):
    """flip -> colour filter -> convert -> rainbow replace: the port against fingerprints of the reference's own
    functions driven in display.c's order (oracle/ref_display_shim.c)"""
    recs, cases = golden["display_frames"], ob.display_cases()
    assert len(recs) == len(cases)
    for rec, case in zip(recs, cases):
        assert all(rec[k] == case[k] for k in case)
        img = ob.gen(case["pattern"], case["W"], case["H"], 0)
        s = ob.port_display_convert(img, **ob.display_args(case))
        got = (None, None) if s is None else (len(s), "%08x" % ob.fnv(s))
        assert got == (rec["bytes"], rec["fnv"]), case


def test_port_vs_ref_display(ob, ref_lib):
    rng = np.random.default_rng(17)
    for it in range(60):
        W, H = int(rng.integers(2, 200)), int(rng.integers(2, 150))
        img = ob.gen(("noise", "bars", "gradient", "grey", "solid")[it % 5], W, H, it)
        kw = dict(cols=int(rng.integers(1, 90)), rows=int(rng.integers(1, 40)), level=int(rng.integers(0, 4)),
                  mode=int(rng.integers(0, 3)), palette=("standard", "blocks")[it % 2], aspect=bool(rng.integers(0, 2)),
                  pad=bool(rng.integers(0, 2)), flip_x=bool(rng.integers(0, 2)), flip_y=bool(rng.integers(0, 2)),
                  color_filter=int(rng.integers(0, 13)), time_s=float(rng.random() * 10))
        assert ob.ref_display_convert(img, **kw) == ob.port_display_convert(img, **kw), kw


def test_port_color_filter_and_rainbow(ob, golden):
    img = ob.gen("noise", 333, 127, 0)
    for rec in golden["color_filter"]:
        rc, out = ob.port_color_filter(img, rec["filter"], rec["time"])
        assert (rc, "%08x" % ob.fnv(out.tobytes())) == (rec["rc"], rec["fnv"]), rec
    for t, r, g, b in golden["rainbow_hue"]:
        assert ob.rainbow_rgb(ob.port().orc_calculate_rainbow, t) == (r, g, b), t
    # invalid arguments, color_filter.c:276-329
    assert ob.port().orc_apply_color_filter(None, 4, 4, 12, 3, 0.0) == -1
    assert ob.port_color_filter(img, 13)[0] == -1 and ob.port_color_filter(img, -1)[0] == -1
    assert ob.port_color_filter(img, 0)[0] == 0


def test_port_vs_ref_color_filter(ob, ref_lib):
    for it, (W, H) in enumerate(((1, 1), (5, 3), (16, 16), (333, 127), (640, 480))):
        img = ob.gen("noise" if it % 2 else "gradient", W, H, it)
        for f in range(-1, 14):
            a, b = ob.ref_color_filter(img, f, 0.37 * it), ob.port_color_filter(img, f, 0.37 * it)
            assert a[0] == b[0] and (a[1] == b[1]).all(), (W, H, f)
    for t in np.linspace(0, 40, 4001):
        assert ob.rainbow_rgb(ref_lib.color_filter_calculate_rainbow, t) == \
            ob.rainbow_rgb(ob.port().orc_calculate_rainbow, t), t


def test_rainbow_replace(ob, ref_lib):
    def both(s, t=1.0):
        a = ob._take(ref_lib.rainbow_replace_ansi_colors(s, t))
        b = ob._take(ob.port().orc_rainbow_replace(s, t))
        assert a == b, s
        return b
    assert both(b"no colour here") is None
    assert both(b"\x1b[48;2;1;2;3mX") is None                      # background SGRs are not touched
    r, g, b = ob.rainbow_rgb(ob.port().orc_calculate_rainbow, 1.0)
    code = b"\x1b[38;2;%d;%d;%dm" % (r, g, b)
    assert both(b"a\x1b[38;2;1;2;3mb\x1b[0m") == b"a" + code + b"b\x1b[0m"
    assert both(b"\x1b[38;2;9;9;9m\x1b[48;2;1;1;1m\xe2\x96\x80") == code + b"\x1b[48;2;1;1;1m\xe2\x96\x80"
    both(b"\x1b[38;2;1;2;3")                                        # unterminated: copied byte by byte
    both(b"\x1b[38;2;\x1b[38;2;4;5;6mz")                            # a start inside a replaced span is swallowed


# ----------------------------------------------------------------- wire packaging (acip/server.c:203-214, crc32.c)
def test_port_crc32c_and_packet_header(ob, golden):
    for rec in golden["crc32c"]:
        if "literal" in rec:
            d = rec["literal"].encode()
            assert "%08x" % ob.port().orc_crc32c(d, len(d)) == rec["crc"] == "e3069283"  # the CRC-32C check value
            continue
        L = rec["len"]
        d = ob.gen("noise", max(1, (L + 2) // 3), 1, 7).tobytes()[:L]
        assert "%08x" % ob.port().orc_crc32c(d, L) == rec["crc"], L
        assert ob.port_packet_header(d, 320, 96).hex() == rec["header"], L


def test_port_vs_ref_crc32c(ob, ref_lib):
    rng = np.random.default_rng(9)
    for L in [0, 1, 2, 7, 8, 9] + [int(v) for v in rng.integers(10, 200000, 40)]:
        d = rng.integers(0, 256, L, dtype=np.uint8).tobytes()
        assert ref_lib.ref_oracle_crc32(d, L) == ref_lib.ref_oracle_crc32_sw(d, L) == ob.port().orc_crc32c(d, L), L
        assert ob.ref_packet_header(d, 203, 61) == ob.port_packet_header(d, 203, 61)


# ----------------------------------------------------------------- 3b. the reference's own KATs for the 8f rows
# tests/unit/video/color_filter_test.c and tests/unit/network/crc32_hw_test.c of the reference checkout
REF_FILTER_COLOURS = {1: (0, 0, 0), 2: (255, 255, 255), 3: (0, 255, 65), 4: (255, 0, 255), 5: (255, 0, 170),
                      6: (255, 136, 0), 7: (0, 221, 221), 8: (0, 255, 255), 9: (255, 182, 193), 10: (255, 51, 51),
                      11: (255, 235, 153)}  # color_filter_test.c:197-212 (metadata_colors)


def check_color_filter_kats(apply):
    """apply(img (h,w,3) uint8, filter, time) -> (rc, filtered); shared by the port (here) and the GPU test"""
    px = np.array([[[0, 0, 0], [255, 255, 255]], [[128, 128, 128], [64, 64, 64]]], np.uint8)
    rc, out = apply(px, 8, 0.0)                                   # colorize_white_on_color (cyan), :95-128
    assert rc == 0 and out[0, 0].tolist() == [0, 0, 0] and out[0, 1].tolist() == [0, 255, 255]
    rc, out = apply(px[:1], 1, 0.0)                               # colorize_black_on_white, :133-152
    assert rc == 0 and all(v < 50 for v in out[0, 0]) and out[0, 1].tolist() == [255, 255, 255]
    one = np.array([[[100, 150, 200]]], np.uint8)
    rc, out = apply(one, 0, 0.0)                                  # apply_none_filter, :157-165
    assert rc == 0 and (out == one).all()
    assert apply(one, 999, 0.0)[0] == -1                          # apply_invalid_params, :190-191
    white = np.full((1, 1, 3), 255, np.uint8)
    for f, rgb in REF_FILTER_COLOURS.items():                     # a white pixel takes the filter's own colour
        if f != 1:
            assert apply(white, f, 0.0)[1][0, 0].tolist() == list(rgb), f


def test_kat_color_filter_reference_unit_tests(ob):
    check_color_filter_kats(ob.port_color_filter)
    # rgb_to_grayscale_* (:19-48) through a white-on-white filter: out = gray * 255 / 255 = gray
    for rgb, lo, hi in (((255, 0, 0), 75, 79), ((0, 255, 0), 148, 152), ((0, 0, 255), 27, 31),
                        ((255, 255, 255), 255, 255), ((0, 0, 0), 0, 0), ((128, 128, 128), 126, 130)):
        g = int(ob.port_color_filter(np.array([[rgb]], np.uint8), 2, 0.0)[1][0, 0, 0])
        assert lo <= g <= hi, (rgb, g)
    assert ob.port().orc_apply_color_filter(None, 1, 1, 3, 3, 0.0) == -1
    buf = (C.c_uint8 * 3)(255, 255, 255)
    for w, h, st in ((0, 1, 3), (1, 0, 3), (1, 1, 0)):            # apply_invalid_params, :174-187
        assert ob.port().orc_apply_color_filter(buf, w, h, st, 3, 0.0) == -1


def test_kat_crc32c_reference_unit_tests(ob):
    crc = lambda b: ob.port().orc_crc32c(b, len(b))  # noqa: E731
    assert crc(b"") == 0                                          # crc32_hw_test.c:14-21
    assert crc(b"Hello, World!") == 0x4D551068                    # :32-50, the one literal the reference pins
    assert crc(b"\x42") != 0 and crc(b"abc") != crc(b"abd") and crc(b"ascii-chat") == crc(b"ascii-chat")


def test_port_vs_ref_dithered_printers(ob, ref_lib):
    """the three Floyd–Steinberg leaf printers (foreground.c:650-749, 752-846 with and without background)"""
    rng = np.random.default_rng(17)
    for it in range(40):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 50))
        img = ob.gen(("noise", "gradient", "bars", "grey")[it % 4], w, h, it)
        for pal in ("standard", "blocks", "minimal"):
            for variant in (0, 1, 2):
                assert ob.port_print_dither(img, pal, variant) == ob.ref_print_dither(img, pal, variant), (it, pal, variant)


def test_box_checker_agrees_with_port(ob, ref_lib):
    """compiled reference's printer on the box-filtered image == the port's box convert"""
    for (W, H, c, r) in ((640, 480, 80, 24), (333, 127, 47, 13), (100, 50, 130, 70)):
        img = ob.gen("noise", W, H, 1)
        for level, mode in ((0, 0), (1, 0), (2, 0), (3, 0), (3, 1), (0, 2), (1, 2), (2, 2), (3, 2)):
            assert ob.ref_box_convert(img, c, r, level, mode) == ob.port_convert(img, c, r, level, mode, scale=ob.SCALE_BOX)


def test_box_fast_equals_box(ob):
    """the column-sum arrangement used by the box-mode CPU baseline is the same function as orc_resize_box"""
    rng = np.random.default_rng(5)
    u8p = C.POINTER(C.c_uint8)
    for it in range(40):
        sw, sh, dw, dh = (int(rng.integers(1, 400)) for _ in range(4))
        if it == 0:
            sw, sh, dw, dh = 3840, 2160, 320, 192
        if it == 1:
            sw, sh, dw, dh = 16, 3000, 4, 5  # tall bands: the u16 guard
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        a = ob.port_resize(src, dw, dh, scale=ob.SCALE_BOX)
        b = np.empty((dh, dw, 3), np.uint8)
        ob.port().orc_resize_box_fast(src.ctypes.data_as(u8p), sw, sh, b.ctypes.data_as(u8p), dw, dh)
        assert np.array_equal(a, b), (sw, sh, dw, dh)


def test_port_vs_ref_digital_rain(ob, ref_lib):
    """digital_rain_apply (lib/video/anim/digital_rain.c:366-520) frame after frame: the state carried between frames,
    the per-visit filter, cursor cells, the rain colour per filter, rainbow mode, malformed strings"""
    n = 0
    for cols, rows, filt, frames in ob.rain_sequences():
        a, b = ob.RefRain(cols, rows, filt), ob.PortRain(cols, rows, filt)
        for i, (s, dt) in enumerate(frames):
            assert a.apply(s, dt) == b.apply(s, dt), (cols, rows, filt, i)
            n += 1
        a.close()
        b.close()
    assert n > 60
