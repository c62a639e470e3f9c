#!/usr/bin/env python
"""Regenerate tests/golden/*.json from the COMPILED REFERENCE (oracle/_ref/libasciichat_ref.so).

Run in the build container only (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures are fingerprints (length, newline count, FNV-1a-32) of the reference's own
output on the deterministic generators of SURVEY.md Appendix C, plus exhaustive-table hashes
of the reference's two colour quantisers and small literal known-answers.  They travel to the
GPU box, where /root/reference does not exist.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_bind as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (pattern, W, H, cols, rows, level, mode, aspect, pad, palette)
FRAME_CASES = [
    # SURVEY.md Appendix C rows (the survey-measured anchors; hashes must come out identical)
    ("noise", 640, 480, 80, 24, 0, 0, 0, 0, "standard"),
    ("noise", 640, 480, 80, 24, 0, 0, 0, 0, "blocks"),
    ("noise", 1920, 1080, 160, 48, 2, 0, 0, 0, "standard"),
    ("noise", 3840, 2160, 320, 96, 3, 2, 0, 0, "standard"),
    ("noise", 3840, 2160, 320, 96, 3, 2, 1, 1, "standard"),
    ("noise", 3840, 2160, 320, 96, 3, 0, 0, 0, "cool"),
    ("noise", 1920, 1080, 160, 48, 1, 0, 0, 0, "standard"),
    ("noise", 1920, 1080, 160, 48, 2, 2, 0, 0, "standard"),
    ("gradient", 1920, 1080, 160, 48, 2, 0, 0, 0, "standard"),
    ("gradient", 3840, 2160, 320, 96, 3, 2, 0, 0, "standard"),
    ("gradient", 3840, 2160, 320, 96, 0, 0, 0, 0, "standard"),
    ("bars", 3840, 2160, 320, 96, 3, 2, 0, 0, "standard"),
    ("bars", 1920, 1080, 160, 48, 3, 0, 1, 1, "standard"),
    ("grey", 1920, 1080, 160, 48, 2, 0, 0, 0, "minimal"),
    ("grey", 3840, 2160, 320, 96, 0, 2, 0, 0, "standard"),
]
# every (level, mode) on every pattern at a small ragged shape and at C1's shape
for pat in ("noise", "gradient", "bars", "grey", "solid"):
    for level in (0, 1, 2, 3):
        for mode in (0, 1, 2):
            FRAME_CASES.append((pat, 333, 127, 47, 13, level, mode, 0, 0, "standard"))
            FRAME_CASES.append((pat, 640, 480, 80, 24, level, mode, 1, 1, "blocks"))
FRAME_CASES += [
    ("noise", 3840, 2160, 320, 96, 3, 0, 0, 0, "standard"),
    ("noise", 1920, 1080, 160, 45, 3, 0, 0, 0, "standard"),
    ("noise", 16, 9, 7, 5, 3, 2, 0, 0, "standard"),  # odd pixel height in half-block (bottom := top)
    ("bars", 64, 64, 64, 32, 3, 2, 0, 0, "standard"),
    ("noise", 7, 3, 1, 1, 0, 0, 0, 0, "standard"),
    ("noise", 1, 1, 1, 1, 3, 0, 0, 0, "standard"),
]

SURVEY_ANCHORS = {  # must be reproduced or the oracle build is wrong
    0: (1943, "0b754fdb"), 1: (1943, "4079e20d"), 2: (88839, "f884e744"), 3: (1180548, "a9e0f66b"),
    4: (1107075, "c06329e7"), 5: (629622, "26225b2d"), 6: (46319, "fbf816a6"), 7: (183725, "2551fac0"),
    8: (87759, "34a09229"), 9: (5190, "c0efc3cc"), 10: (767, "5c5d8d91"), 11: (32991, "193e302d"),
    12: (17105, "00a0cef4"), 13: (92399, "71779d6d"), 14: (88031, "46c5e26d"),
}


def main():
    R = ob.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    frames = []
    for i, (pat, W, H, c, r, level, mode, aspect, pad, pal) in enumerate(FRAME_CASES):
        img = ob.gen(pat, W, H, 0)
        s = ob.ref_convert(img, c, r, level, mode, pal, bool(aspect), False, bool(pad))
        rec = dict(pattern=pat, W=W, H=H, cols=c, rows=r, level=level, mode=mode, aspect=aspect, pad=pad,
                   palette=pal, bytes=len(s), newlines=s.count(b"\n"), fnv="%08x" % ob.fnv(s))
        if i in SURVEY_ANCHORS:
            assert (rec["bytes"], rec["fnv"]) == SURVEY_ANCHORS[i], (i, rec)
        frames.append(rec)

    # exhaustive quantiser tables from the reference's own functions
    tab = np.empty(1 << 24, np.uint8)
    u8p = C.POINTER(C.c_uint8)
    ob.port().orc_fill_table(2, C.cast(R.rgb_to_256color, C.c_void_p), tab.ctypes.data_as(u8p))
    h256 = "%08x" % ob.fnv(tab.tobytes())
    ob.port().orc_fill_table(2, C.cast(R.rgb_to_16color, C.c_void_p), tab.ctypes.data_as(u8p))
    h16 = "%08x" % ob.fnv(tab.tobytes())

    # the quirk literals of SURVEY.md §8a (all-white 8x2, standard palette)
    white = np.full((2, 8, 3), 255, np.uint8)
    quirks = {}
    for name, level, mode in (("Q1_mono", 0, 0), ("Q2_16", 1, 0), ("Q3_256", 2, 0), ("Q3_true", 3, 0),
                              ("Q4_true_bg", 3, 1), ("Q5_half", 3, 2)):
        quirks[name] = ob.ref_print(white, level, mode, "standard").decode("latin-1")

    # NN index maths: sampled source coordinates for awkward ratios
    nn = []
    for (sw, sh, dw, dh) in ((3840, 2160, 320, 192), (1920, 1080, 160, 48), (640, 480, 80, 24), (333, 127, 47, 13),
                             (5, 3, 11, 7), (10000, 1, 3, 1), (1, 1, 4, 4), (320, 240, 320, 240)):
        src = np.zeros((sh, sw, 3), np.uint8)
        xs = np.arange(sw, dtype=np.uint32)
        ys = np.arange(sh, dtype=np.uint32)
        src[:, :, 0] = (xs & 255)[None, :]
        src[:, :, 1] = ((xs >> 8) & 255)[None, :] | ((ys[:, None] >> 8) << 6).astype(np.uint8)
        src[:, :, 2] = (ys & 255)[:, None]
        out = ob.ref_resize(src, dw, dh)
        nn.append(dict(sw=sw, sh=sh, dw=dw, dh=dh, fnv="%08x" % ob.fnv(out.tobytes())))

    # glyph tables as the reference's cache builds them, observed through 1-row renders
    # (grey ramp 0..255 rendered in each mode exposes cache[Y], Q1 and Q2 mappings)
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)
    glyphs = {}
    for pal in ob.PALETTES:
        glyphs[pal] = {
            "mono": "%08x" % ob.fnv(ob.ref_print(ramp, 0, 0, pal)),
            "c16": "%08x" % ob.fnv(ob.ref_print(ramp, 1, 0, pal)),
            "c256": "%08x" % ob.fnv(ob.ref_print(ramp, 2, 0, pal)),
            "true": "%08x" % ob.fnv(ob.ref_print(ramp, 3, 0, pal)),
        }

    # text-space grid compositor
    grids = []
    for n, cols, rows, W, H, level, mode in ((1, 40, 12, 80, 24, 0, 0), (2, 40, 12, 80, 24, 2, 0),
                                             (3, 40, 12, 120, 40, 3, 0), (4, 30, 10, 100, 30, 0, 0),
                                             (8, 40, 12, 160, 48, 3, 2), (9, 20, 6, 80, 24, 1, 0),
                                             (2, 40, 12, 15, 5, 0, 0)):
        srcs = [ob.ref_convert(ob.gen("noise" if i % 2 else "bars", 160, 120, i), cols, rows, level, mode)
                for i in range(n)]
        g, sz = ob.ref_create_grid(srcs, W, H)
        grids.append(dict(n=n, cols=cols, rows=rows, W=W, H=H, level=level, mode=mode, size=sz,
                          fnv="%08x" % ob.fnv(g)))

    # server pixel-space compositor (stream.c:523-779 compiled through oracle/ref_stream_shim.c)
    comps = []
    rng = np.random.default_rng(5)
    for it in range(30):
        n = int(rng.integers(1, 10))
        pats = [("noise", "bars", "gradient")[i % 3] for i in range(n)]
        dims = [(int(rng.integers(40, 400)), int(rng.integers(30, 300))) for _ in range(n)]
        W, H = int(rng.integers(40, 200)), int(rng.integers(20, 60))
        srcs = [ob.gen(p, w, h, i) for i, (p, (w, h)) in enumerate(zip(pats, dims))]
        out, c, r = ob.ref_composite(srcs, W, H)
        comps.append(dict(n=n, W=W, H=H, cols=c, rows=r, fnv="%08x" % ob.fnv(out.tobytes())))

    # the server's whole per-client entry create_mixed_ascii_frame_for_client (stream.c:958-1191), run inside
    # the compiled reference by oracle/ref_stream_shim.c's synthetic clients
    mixed = []
    for case in ob.mixed_cases():
        srcs = ob.mixed_sources(case)
        s, sz, cnt = ob.ref_mixed_frame(srcs, case["W"], case["H"], case["level"], case["mode"], case["palette"],
                                        bool(case["pad"]))
        mixed.append(dict(case, size=sz, sources=cnt, fnv=None if s is None else "%08x" % ob.fnv(s)))

    # client display path (display.c:484-671: flip -> apply_color_filter -> convert -> rainbow replace), every byte
    # produced by reference code driven by oracle/ref_display_shim.c
    display = []
    for case in ob.display_cases():
        img = ob.gen(case["pattern"], case["W"], case["H"], 0)
        s = ob.ref_display_convert(img, **ob.display_args(case))
        display.append(dict(case, bytes=None if s is None else len(s), fnv=None if s is None else "%08x" % ob.fnv(s)))
    # whole-image colour filter (color_filter.c:274-346) and the rainbow hue (:165-236)
    filt = []
    img = ob.gen("noise", 333, 127, 0)
    for f in range(13):
        rc, out = ob.ref_color_filter(img, f, 1.7)
        filt.append(dict(filter=f, time=1.7, rc=rc, fnv="%08x" % ob.fnv(out.tobytes())))
    rainbow = [[round(float(t), 3)] + list(ob.rainbow_rgb(R.color_filter_calculate_rainbow, round(float(t), 3)))
               for t in np.linspace(0.0, 9.0, 61)]
    # wire packaging (acip/server.c:203-214 header, crc32.c): the reference's own CRC32-C and header bytes
    crc = []
    for L in (0, 1, 3, 63, 64, 65, 4095, 16384, 16385, 100001, 1180548):
        d = ob.gen("noise", max(1, (L + 2) // 3), 1, 7).tobytes()[:L]
        crc.append(dict(len=L, crc="%08x" % R.ref_oracle_crc32(d, L), header=ob.ref_packet_header(d, 320, 96).hex()))
    crc.append(dict(literal="123456789", crc="%08x" % R.ref_oracle_crc32(b"123456789", 9)))

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py",
                       reference_commit="73fe49337008f687add06622c012ac1df0ab3dcc",
                       frames=frames, rgb_to_256color_table_fnv=h256, rgb_to_16color_table_fnv=h16,
                       quirks=quirks, nn_resize=nn, glyph_tables=glyphs, text_grids=grids, pixel_composites=comps,
                       mixed_frames=mixed, display_frames=display, color_filter=filt, rainbow_hue=rainbow,
                       crc32c=crc), f, indent=1)
    print("wrote", len(frames), "frame fingerprints;", "q256", h256, "q16", h16)


if __name__ == "__main__":
    main()
