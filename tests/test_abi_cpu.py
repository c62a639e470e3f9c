"""CPU-only checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/asciichat_b200.h declares, its struct layouts equal the reference's, and — with no GPU —
it fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def acb():
    import ascii_chat_b200 as m
    m.build_library()
    return m


def test_library_exports_every_declared_symbol(acb):
    hdr = open(os.path.join(ROOT, "include", "asciichat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", hdr)) - {"__attribute__", "push", "defined", "packed", "void"}
    assert declared == set(acb.binding.EXPORTS), declared ^ set(acb.binding.EXPORTS)
    L = acb.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_library_is_sm100a_and_self_contained(acb):
    out = subprocess.run(["cuobjdump", "--list-elf", acb.binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    needed = subprocess.run(["readelf", "-d", acb.binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" not in needed  # static cudart: only the driver is needed at run time
    assert "oracle" not in needed and "torch" not in needed


def test_struct_layouts_match_reference(acb, ob):
    assert C.sizeof(acb.terminal_capabilities_t) == C.sizeof(ob.Caps) == 240
    assert C.sizeof(acb.image_t) == C.sizeof(ob.Image) == 24
    for f in ("color_level", "render_mode", "wants_padding", "pad_height", "palette_custom"):
        assert getattr(acb.terminal_capabilities_t, f).offset == getattr(ob.Caps, f).offset
    # offsets as the reference's own header lays them out (checked with the reference headers when present)
    ref_inc = "/root/reference/include"
    if os.path.isdir(ref_inc):
        src = r'''
#include <stddef.h>
#include <stdio.h>
#include <ascii-chat/platform/terminal.h>
#include <ascii-chat/video/rgba/image.h>
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(terminal_capabilities_t),
 offsetof(terminal_capabilities_t, render_mode), offsetof(terminal_capabilities_t, wants_padding),
 offsetof(terminal_capabilities_t, pad_height), sizeof(image_t), offsetof(image_t, pixels));return 0;}'''
        exe = "/tmp/acb200_layout_probe"
        inc = ["-I", os.path.join(ROOT, "oracle", "_ref", "shim_inc"), "-I", ref_inc, "-I", "/root/reference/deps",
               "-I", "/root/reference/deps/ascii-chat-deps/uthash/src"]
        r = subprocess.run(["gcc", "-std=gnu2x", "-w", "-x", "c", "-", "-o", exe] + inc, input=src, text=True,
                           capture_output=True)
        if r.returncode == 0:
            vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
            T = acb.terminal_capabilities_t
            assert vals == [C.sizeof(T), T.render_mode.offset, T.wants_padding.offset, T.pad_height.offset,
                            C.sizeof(acb.image_t), acb.image_t.pixels.offset]


def test_host_float_helpers_match_oracle(acb, ob):
    rng = np.random.default_rng(1)
    for _ in range(2000):
        iw, ih, w, h = (int(rng.integers(1, 4000)) for _ in range(4))
        ow, oh = C.c_long(0), C.c_long(0)
        ob.port().orc_aspect_ratio(iw, ih, w, h, 0, C.byref(ow), C.byref(oh))
        assert acb.aspect_ratio(iw, ih, w, h) == (ow.value, oh.value)
    for _ in range(300):
        n = int(rng.integers(1, 10))
        ws = [int(rng.integers(16, 2000)) for _ in range(n)]
        hs = [int(rng.integers(16, 2000)) for _ in range(n)]
        tw, th = int(rng.integers(20, 400)), int(rng.integers(10, 120))
        a, b = C.c_int(0), C.c_int(0)
        c, d = C.c_int(0), C.c_int(0)
        arr = lambda v: (C.c_int * n)(*v)  # noqa: E731
        ob.port().orc_grid_layout(arr(ws), arr(hs), n, tw, th, C.byref(a), C.byref(b))
        acb.lib().acb200_grid_layout(arr(ws), arr(hs), n, tw, th, C.byref(c), C.byref(d))
        assert (a.value, b.value) == (c.value, d.value)


def test_nn_row_schedule_matches_image_resize_indices(acb):
    """the row list behind the copy-engine fetch (engine.cu: row_schedule) == image.c:294,315-317's sy(y), mirrored and
    walked backwards with flip_y, whenever a decomposition into <= 8 arithmetic progressions is reported"""
    import ctypes as C
    import random
    rnd = random.Random(3)
    cases = [(2160, 192), (2160, 96), (2160, 180), (1080, 96), (1080, 48), (480, 48), (480, 24), (720, 134), (5, 3),
             (1, 1), (7, 7), (3, 9), (2160, 1), (2160, 2)] + [(rnd.randint(1, 2200), rnd.randint(1, 400)) for _ in range(300)]
    found = 0
    for (sh, rows), flip in [(c, f) for c in cases for f in (0, 1)]:
        P, D, first = C.c_int(), C.c_int(), (C.c_int * 8)()
        assert acb.lib().acb200_nn_row_schedule(sh, rows, flip, C.byref(P), C.byref(D), first) == 0
        yr = ((sh << 16) // rows) + 1
        sy = [min(((y * yr) & 0xFFFFFFFF) >> 16, sh - 1) for y in range(rows)]
        want = sy if not flip else [sh - 1 - v for v in reversed(sy)]  # ascending list; row j is output row rows-1-j
        if P.value == 0:
            continue
        found += 1
        assert 1 <= P.value <= 8
        got = [first[j % P.value] + (j // P.value) * D.value for j in range(rows)]
        assert got == want, (sh, rows, flip, P.value, D.value)
    assert found >= 40  # the BASELINE geometries and every integer ratio have one
    for sh, rows in ((2160, 192), (2160, 96), (1080, 96), (480, 48)):
        P, D, first = C.c_int(), C.c_int(), (C.c_int * 8)()
        acb.lib().acb200_nn_row_schedule(sh, rows, 0, C.byref(P), C.byref(D), first)
        assert P.value in (1, 2, 4), (sh, rows, P.value)


def test_no_gpu_means_loud_failure(acb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    caps = acb.make_caps(3, 2)
    assert acb.ascii_convert_with_capabilities(np.zeros((8, 8, 3), np.uint8), 4, 2, caps, False, False,
                                               "standard") is None
    code, msg = acb.last_error()
    assert code == 85 and "no CPU path" in msg
    with pytest.raises(RuntimeError):
        acb.render_batch_host(acb.make_cfg(8, 8, 4, 4, 3, 2), [np.zeros((8, 8, 3), np.uint8)])


def test_wire_ingest_validation_needs_no_gpu(acb):
    """the IMAGE_FRAME payload checks of handle_image_frame_packet (protocol.c:748-803) run before any device work"""
    import struct
    ok_hdr = struct.pack(">II", 4, 2)
    for payload in (b"", b"\0" * 7, struct.pack(">II", 0, 2) + b"x" * 24, struct.pack(">II", 3841, 1) + b"x" * 11523,
                    struct.pack(">II", 1, 2161) + b"x" * 6483, ok_hdr + b"x" * 23, ok_hdr + b"x" * 25):
        assert acb.source_update_wire(0, payload) == 86, payload[:8]
        assert acb.last_error()[0] == 86
    assert acb.source_update_wire(99, ok_hdr + b"x" * 24) == 86


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "ascii-chat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_bind" not in txt and "liboracle" not in txt and "libasciichat_ref" not in txt, f
