"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``port()``  -> oracle/liboracle_port.so   : our CPU restatement (oracle/ascii_oracle.c)
* ``ref()``   -> oracle/_ref/libasciichat_ref.so : the reference's own sources compiled
                 unmodified by oracle/Makefile (None when it was never built)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  Nothing under ascii-chat_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle_port.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libasciichat_ref.so")
REFERENCE_ROOT = "/root/reference"

PALETTES = {  # include/ascii-chat/video/ascii/palette.h:161-197
    "standard": "   ...',;:clodxkO0KXNWM",
    "blocks": "   ░░▒▒▓▓██",
    "digital": "   -=≡≣▰▱◼",
    "minimal": "   .-+*#",
    "cool": "   ▁▂▃▄▅▆▇█",
}
PATTERNS = {"noise": 0, "gradient": 1, "bars": 2, "grey": 3, "solid": 4}

COLOR_NONE, COLOR_16, COLOR_256, COLOR_TRUE = 0, 1, 2, 3
MODE_FG, MODE_BG, MODE_HALF = 0, 1, 2
SCALE_NN, SCALE_BOX = 0, 1


class Caps(C.Structure):
    """terminal_capabilities_t — include/ascii-chat/platform/terminal.h:707-738 (240 bytes on LP64)."""

    _fields_ = [
        ("color_level", C.c_int), ("capabilities", C.c_uint32), ("color_count", C.c_uint32),
        ("utf8_support", C.c_bool), ("detection_reliable", C.c_bool), ("render_mode", C.c_int),
        ("term_type", C.c_char * 64), ("colorterm", C.c_char * 64), ("wants_background", C.c_bool),
        ("palette_type", C.c_int), ("palette_custom", C.c_char * 64), ("desired_fps", C.c_uint8),
        ("color_filter", C.c_int), ("wants_padding", C.c_bool), ("pad_height", C.c_size_t),
    ]


class Image(C.Structure):
    """image_t — include/ascii-chat/video/rgba/image.h:143-148."""

    _fields_ = [("w", C.c_int), ("h", C.c_int), ("pixels", C.c_void_p), ("alloc_method", C.c_uint8)]


class FrameSource(C.Structure):
    """ascii_frame_source_t — include/ascii-chat/video/ascii/ascii.h:358-361."""

    _fields_ = [("frame_data", C.c_char_p), ("frame_size", C.c_size_t)]


def make_caps(level, mode, pad=False):
    c = Caps()
    c.color_level, c.render_mode, c.utf8_support, c.wants_padding = level, mode, True, pad
    return c


def _make(target):
    subprocess.run(["make", "-C", ORACLE_DIR, target], check=True, stdout=subprocess.DEVNULL)


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _take(ptr, n=None):
    """copy a malloc'd C string into bytes and free it"""
    if not ptr:
        return None
    out = C.string_at(ptr) if n is None else C.string_at(ptr, n)
    _libc.free(ptr)
    return out


def as_u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


_port = None


def port():
    global _port
    if _port is not None:
        return _port
    src = os.path.join(ORACLE_DIR, "ascii_oracle.c")
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        _make("port")
    L = C.CDLL(PORT_SO)
    u8p = C.POINTER(C.c_uint8)
    L.orc_print.restype = C.c_void_p
    L.orc_print.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_size_t)]
    L.orc_convert_caps.restype = C.c_void_p
    L.orc_convert_caps.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_size_t)]
    L.orc_convert.restype = C.c_void_p
    L.orc_convert.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_char_p,
                              C.c_int, C.POINTER(C.c_size_t)]
    L.orc_pad_width.restype = C.c_void_p
    L.orc_pad_width.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_pad_height.restype = C.c_void_p
    L.orc_pad_height.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_resize_nn.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.orc_resize_box.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.orc_resize_box_fast.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.orc_bench_box.restype = C.c_double
    L.orc_bench_box.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_gen_pattern.argtypes = [C.c_int, C.c_uint32, u8p, C.c_int, C.c_int]
    L.orc_fnv1a32.restype = C.c_uint32
    L.orc_fnv1a32.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_fill_table.argtypes = [C.c_int, C.c_void_p, u8p]
    L.orc_rgb_to_256.argtypes = [C.c_int] * 3
    L.orc_rgb_to_16.argtypes = [C.c_int] * 3
    L.orc_rep_is_profitable.argtypes = [C.c_uint32]
    L.orc_digits_u32.argtypes = [C.c_uint32]
    L.orc_aspect_ratio.argtypes = [C.c_long] * 4 + [C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.orc_build_glyph_lut.argtypes = [C.c_char_p, C.c_int, u8p, u8p]
    L.orc_create_grid.restype = C.c_void_p
    L.orc_create_grid.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_size_t)]
    L.orc_grid_layout.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_composite.argtypes = [C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_bench_convert.restype = C.c_double
    L.orc_bench_convert.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_char_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_calculate_rainbow.restype = None
    L.orc_calculate_rainbow.argtypes = [C.c_float, u8p, u8p, u8p]
    L.orc_apply_color_filter.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float]
    L.orc_rainbow_replace.restype = C.c_void_p
    L.orc_rainbow_replace.argtypes = [C.c_char_p, C.c_float]
    L.orc_rain_init.restype = C.c_void_p
    L.orc_rain_init.argtypes = [C.c_int, C.c_int]
    L.orc_rain_destroy.restype = None
    L.orc_rain_destroy.argtypes = [C.c_void_p]
    L.orc_rain_set_filter.restype = None
    L.orc_rain_set_filter.argtypes = [C.c_void_p, C.c_int]
    L.orc_rain_apply.restype = C.c_void_p
    L.orc_rain_apply.argtypes = [C.c_void_p, C.c_char_p, C.c_float, C.POINTER(C.c_size_t)]
    L.orc_print_dither.restype = C.c_void_p
    L.orc_print_dither.argtypes = [u8p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_size_t)]
    L.orc_display_convert.restype = C.c_void_p
    L.orc_display_convert.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                      C.POINTER(C.c_size_t)]
    L.orc_crc32c.restype = C.c_uint32
    L.orc_crc32c.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_frame_packet_header.restype = None
    L.orc_frame_packet_header.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, u8p]
    _port = L
    return L


_ref = False


def ref():
    """the compiled reference, or None if oracle/_ref was never built and cannot be built here"""
    global _ref
    if _ref is not False:
        return _ref
    if not os.path.exists(REF_SO) and os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "video", "ascii")):
        _make("ref")
    if not os.path.exists(REF_SO):
        _ref = None
        return None
    L = C.CDLL(REF_SO)
    ip, cp = C.POINTER(Image), C.POINTER(Caps)
    L.ascii_convert_with_capabilities.restype = C.c_void_p
    L.ascii_convert_with_capabilities.argtypes = [ip, C.c_ssize_t, C.c_ssize_t, cp, C.c_bool, C.c_bool, C.c_char_p]
    L.ascii_convert.restype = C.c_void_p
    L.ascii_convert.argtypes = [ip, C.c_ssize_t, C.c_ssize_t, C.c_bool, C.c_bool, C.c_bool, C.c_char_p, C.c_char_p]
    L.image_print_with_capabilities.restype = C.c_void_p
    L.image_print_with_capabilities.argtypes = [ip, cp, C.c_char_p]
    L.image_resize.argtypes = [ip, ip]
    L.ascii_create_grid.restype = C.c_void_p
    L.ascii_create_grid.argtypes = [C.POINTER(FrameSource), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.ascii_pad_frame_width.restype = C.c_void_p
    L.ascii_pad_frame_width.argtypes = [C.c_char_p, C.c_size_t]
    L.ascii_pad_frame_height.restype = C.c_void_p
    L.ascii_pad_frame_height.argtypes = [C.c_char_p, C.c_size_t]
    L.rgb_to_256color.restype = C.c_uint8
    L.rgb_to_256color.argtypes = [C.c_uint8] * 3
    L.rgb_to_16color.restype = C.c_uint8
    L.rgb_to_16color.argtypes = [C.c_uint8] * 3
    L.rep_is_profitable.restype = C.c_bool
    L.rep_is_profitable.argtypes = [C.c_uint32]
    L.aspect_ratio.argtypes = [C.c_ssize_t] * 4 + [C.c_bool, C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t)]
    L.append_truecolor_fg.restype = C.c_void_p
    L.append_truecolor_fg.argtypes = [C.c_char_p, C.c_uint8, C.c_uint8, C.c_uint8]
    L.append_truecolor_bg.restype = C.c_void_p
    L.append_truecolor_bg.argtypes = [C.c_char_p, C.c_uint8, C.c_uint8, C.c_uint8]
    L.append_256color_fg.restype = C.c_void_p
    L.append_256color_fg.argtypes = [C.c_char_p, C.c_uint8]
    L.append_16color_fg.restype = C.c_void_p
    L.append_16color_fg.argtypes = [C.c_char_p, C.c_uint8]
    L.append_16color_bg.restype = C.c_void_p
    L.append_16color_bg.argtypes = [C.c_char_p, C.c_uint8]
    L.ansi_fast_init_256color.restype = None
    L.ansi_fast_init_16color.restype = None
    L.ascii_simd_init.restype = None
    L.ref_oracle_set_render_mode.argtypes = [C.c_int]
    u8pp = C.POINTER(C.POINTER(C.c_uint8))
    L.ref_oracle_composite.argtypes = [u8pp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_uint8)]
    L.ref_oracle_grid_layout.restype = None
    L.ref_oracle_grid_layout.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
    u8p = C.POINTER(C.c_uint8)
    L.apply_color_filter.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float]
    L.color_filter_calculate_rainbow.restype = None
    L.color_filter_calculate_rainbow.argtypes = [C.c_float, u8p, u8p, u8p]
    L.rainbow_replace_ansi_colors.restype = C.c_void_p
    L.rainbow_replace_ansi_colors.argtypes = [C.c_char_p, C.c_float]
    L.image_print_16color_dithered.restype = C.c_void_p
    L.image_print_16color_dithered.argtypes = [ip, C.c_char_p]
    L.image_print_16color_dithered_with_background.restype = C.c_void_p
    L.image_print_16color_dithered_with_background.argtypes = [ip, C.c_bool, C.c_char_p]
    L.digital_rain_init.restype = C.c_void_p
    L.digital_rain_init.argtypes = [C.c_int, C.c_int]
    L.digital_rain_destroy.restype = None
    L.digital_rain_destroy.argtypes = [C.c_void_p]
    L.digital_rain_apply.restype = C.c_void_p
    L.digital_rain_apply.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    L.digital_rain_set_color_from_filter.restype = None
    L.digital_rain_set_color_from_filter.argtypes = [C.c_void_p, C.c_int]
    L.digital_rain_reset.restype = None
    L.digital_rain_reset.argtypes = [C.c_void_p]
    L.ref_oracle_display_convert.restype = C.c_void_p
    L.ref_oracle_display_convert.argtypes = [u8p, C.c_int, C.c_int, C.c_long, C.c_long, cp, C.c_int, C.c_int,
                                             C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_size_t)]
    L.ref_oracle_frame_packet_header.restype = None
    L.ref_oracle_frame_packet_header.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, u8p]
    L.ref_oracle_crc32.restype = C.c_uint32
    L.ref_oracle_crc32.argtypes = [C.c_char_p, C.c_size_t]
    L.ref_oracle_crc32_sw.restype = C.c_uint32
    L.ref_oracle_crc32_sw.argtypes = [C.c_char_p, C.c_size_t]
    L.ascii_simd_init()
    L.ansi_fast_init_256color()
    L.ansi_fast_init_16color()
    _ref = L
    return L


# ----------------------------------------------------------------------------- helpers
def gen(kind, w, h, frame=0):
    """synthetic RGB24 frame (SURVEY.md Appendix C generators), shape (h, w, 3) uint8"""
    out = np.empty((h, w, 3), np.uint8)
    port().orc_gen_pattern(PATTERNS[kind] if isinstance(kind, str) else kind, frame,
                           out.ctypes.data_as(C.POINTER(C.c_uint8)), w, h)
    return out


def fnv(b):
    return int(port().orc_fnv1a32(b, len(b)))


def pal_bytes(p):
    return PALETTES.get(p, p).encode() if isinstance(p, str) else p


def port_convert(img, cols, rows, level, mode, palette="standard", aspect=False, stretch=False, pad=False,
                 scale=SCALE_NN):
    a, p = as_u8(img)
    n = C.c_size_t(0)
    r = port().orc_convert_caps(p, a.shape[1], a.shape[0], cols, rows, level, mode, int(pad), int(aspect),
                                int(stretch), pal_bytes(palette), scale, C.byref(n))
    return _take(r)


def port_print(img, level, mode, palette="standard"):
    a, p = as_u8(img)
    n = C.c_size_t(0)
    return _take(port().orc_print(p, a.shape[1], a.shape[0], level, mode, pal_bytes(palette), C.byref(n)))


def port_convert_legacy(img, cols, rows, color, aspect, stretch, palette="standard", opt_mode=0):
    a, p = as_u8(img)
    n = C.c_size_t(0)
    return _take(port().orc_convert(p, a.shape[1], a.shape[0], cols, rows, int(color), int(aspect), int(stretch),
                                    pal_bytes(palette), opt_mode, C.byref(n)))


def port_resize(img, dw, dh, scale=SCALE_NN):
    a, p = as_u8(img)
    out = np.empty((dh, dw, 3), np.uint8)
    fn = port().orc_resize_box if scale == SCALE_BOX else port().orc_resize_nn
    fn(p, a.shape[1], a.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)), dw, dh)
    return out


def ref_convert(img, cols, rows, level, mode, palette="standard", aspect=False, stretch=False, pad=False):
    a, _ = as_u8(img)
    im = Image(a.shape[1], a.shape[0], a.ctypes.data, 0)
    caps = make_caps(level, mode, pad)
    return _take(ref().ascii_convert_with_capabilities(C.byref(im), cols, rows, C.byref(caps), aspect, stretch,
                                                       pal_bytes(palette)))


def ref_print(img, level, mode, palette="standard"):
    a, _ = as_u8(img)
    im = Image(a.shape[1], a.shape[0], a.ctypes.data, 0)
    caps = make_caps(level, mode)
    return _take(ref().image_print_with_capabilities(C.byref(im), C.byref(caps), pal_bytes(palette)))


def port_print_dither(img, palette="standard", variant=0):
    """variant 0 = ..._with_background(img, true), 1 = ..._with_background(img, false), 2 = image_print_16color_dithered"""
    a, p = as_u8(img)
    n = C.c_size_t(0)
    return _take(port().orc_print_dither(p, a.shape[1], a.shape[0], pal_bytes(palette), variant, C.byref(n)))


def ref_print_dither(img, palette="standard", variant=0):
    a, _ = as_u8(img)
    im = Image(a.shape[1], a.shape[0], a.ctypes.data, 0)
    if variant == 2:
        return _take(ref().image_print_16color_dithered(C.byref(im), pal_bytes(palette)))
    return _take(ref().image_print_16color_dithered_with_background(C.byref(im), variant == 0, pal_bytes(palette)))


class RefRain:
    """the compiled reference's digital_rain_t (lib/video/anim/digital_rain.c), driven frame by frame"""

    def __init__(self, cols, rows, color_filter=0):
        self.p = ref().digital_rain_init(cols, rows)
        ref().digital_rain_set_color_from_filter(self.p, int(color_filter))

    def apply(self, frame, dt):
        return _take(ref().digital_rain_apply(self.p, frame, float(dt)))

    def reset(self):
        ref().digital_rain_reset(self.p)

    def close(self):
        ref().digital_rain_destroy(self.p)


class PortRain:
    def __init__(self, cols, rows, color_filter=0):
        self.p = port().orc_rain_init(cols, rows)
        port().orc_rain_set_filter(self.p, int(color_filter))

    def apply(self, frame, dt):
        n = C.c_size_t(0)
        return _take(port().orc_rain_apply(self.p, frame, float(dt), C.byref(n)))

    def close(self):
        port().orc_rain_destroy(self.p)


def rain_sequences():
    """deterministic multi-frame inputs for the digital-rain checks: (cols, rows, filter, [(frame string, dt), ...])"""
    out = []
    for k, (cols, rows, level, mode, filt, pal) in enumerate(((80, 24, 3, 0, 0, "standard"), (80, 24, 0, 0, 3, "standard"),
                                                           (160, 48, 3, 2, 12, "standard"), (40, 12, 2, 0, 7, "blocks"),
                                                           (100, 30, 3, 0, 1, "cool"), (64, 20, 1, 2, 0, "standard"))):
        frames = []
        for f in range(8):
            s = port_convert(gen(("noise", "bars", "gradient")[(f + k) % 3], 320, 240, f), cols, rows, level, mode, pal)
            if f == 5:
                s = s + b"\n\n" + s[:300] + b"\n"      # more rows than the grid, trailing newline
            if f == 6:
                s = s[: len(s) // 2]                       # cut mid-frame (fewer rows, maybe mid-sequence)
            frames.append((s, 0.033 * (1 + f % 3)))
        out.append((cols, rows, filt, frames))
    # hand-made strings: other escape sequences, lone ESC, ESC before newline, invalid UTF-8, a CSI that swallows a
    # newline (the serial fallback), bg+fg pairs on one cell, colours beyond 255, an empty frame
    e = b"\x1b"
    odd = [b"plain\ntext", e + b"[0m" + b"A" + e + b"[7bB\n" + e + b"\n" + e, b"\xe2\x96\x80\xe2\x96x\xff\xc3(\n\xf0\x9f\x98\x80!",
           e + b"[38;2;300;2;1mX" + e + b"[48;2;9;9;9m" + e + b"[38;2;1;2;3mY\n" + e + b"[38;2;1;2mZ" + e + b"[38;5;7mQ",
           b"ab" + e + b"[12\n34mcd\nef", b"", b"\n\n\n", b"x" * 300 + b"\n" + b"y" * 5]
    out.append((10, 3, 5, [(s, 0.05) for s in odd] + [(s, 0.1) for s in odd]))
    return out


def ref_rainbow_replace(s, t):
    return _take(ref().rainbow_replace_ansi_colors(s, float(t)))


def port_rainbow_replace(s, t):
    return _take(port().orc_rainbow_replace(s, float(t)))


def ref_box_convert(img, cols, rows, level, mode, palette="standard"):
    """box mode's checker is the COMPILED reference's printer fed the box-filtered image (the box filter
    itself is this repo's specification, orc_resize_box; the reference has none)"""
    a, _ = as_u8(img)
    rows_px = rows * 2 if mode == 2 else rows
    small = port_resize(a, cols, rows_px, scale=SCALE_BOX)
    return ref_print(small, level, mode, palette)


def ref_convert_legacy(img, cols, rows, color, aspect, stretch, palette="standard", opt_mode=0):
    a, _ = as_u8(img)
    im = Image(a.shape[1], a.shape[0], a.ctypes.data, 0)
    ref().ref_oracle_set_render_mode(opt_mode)
    lum = bytes(range(1, 256)) + b"\x01"
    r = _take(ref().ascii_convert(C.byref(im), cols, rows, color, aspect, stretch, pal_bytes(palette), lum))
    ref().ref_oracle_set_render_mode(0)
    return r


def ref_resize(img, dw, dh):
    a, _ = as_u8(img)
    out = np.zeros((dh, dw, 3), np.uint8)
    s = Image(a.shape[1], a.shape[0], a.ctypes.data, 0)
    d = Image(dw, dh, out.ctypes.data, 0)
    ref().image_resize(C.byref(s), C.byref(d))
    return out


def ref_create_grid(frames, width, height):
    arr = (FrameSource * len(frames))()
    keep = []
    for i, f in enumerate(frames):
        if f is None:
            arr[i].frame_data, arr[i].frame_size = None, 0
        else:
            keep.append(f)
            arr[i].frame_data, arr[i].frame_size = f, len(f)
    n = C.c_size_t(0)
    r = ref().ascii_create_grid(arr, len(frames), width, height, C.byref(n))
    return _take(r), n.value


def port_create_grid(frames, width, height):
    k = len(frames)
    ptrs = (C.c_char_p * k)(*[f for f in frames])
    sizes = (C.c_size_t * k)(*[0 if f is None else len(f) for f in frames])
    n = C.c_size_t(0)
    r = port().orc_create_grid(ptrs, sizes, k, width, height, C.byref(n))
    return _take(r), n.value


def port_composite(srcs, width, height):
    k = len(srcs)
    arrs = [np.ascontiguousarray(s, np.uint8) for s in srcs]
    u8p = C.POINTER(C.c_uint8)
    ptrs = (u8p * k)(*[a.ctypes.data_as(u8p) for a in arrs])
    ws = (C.c_int * k)(*[a.shape[1] for a in arrs])
    hs = (C.c_int * k)(*[a.shape[0] for a in arrs])
    out = np.empty((height * 2, width, 3), np.uint8)
    c, r = C.c_int(0), C.c_int(0)
    port().orc_composite(ptrs, ws, hs, k, width, height, out.ctypes.data_as(u8p), C.byref(c), C.byref(r))
    return out, c.value, r.value


def ref_composite(srcs, width, height):
    """the reference's own create_multi_source_composite / calculate_optimal_grid_layout (src/server/stream.c:523-779),
    compiled into oracle/_ref through oracle/ref_stream_shim.c"""
    k = len(srcs)
    arrs = [np.ascontiguousarray(s, np.uint8) for s in srcs]
    u8p = C.POINTER(C.c_uint8)
    ptrs = (u8p * k)(*[a.ctypes.data_as(u8p) for a in arrs])
    ws = (C.c_int * k)(*[a.shape[1] for a in arrs])
    hs = (C.c_int * k)(*[a.shape[0] for a in arrs])
    out = np.empty((height * 2, width, 3), np.uint8)
    rc = ref().ref_oracle_composite(ptrs, ws, hs, k, width, height, out.ctypes.data_as(u8p))
    assert rc == 0
    c, r = C.c_int(0), C.c_int(0)
    ref().ref_oracle_grid_layout(ws, hs, k, width, height, C.byref(c), C.byref(r))
    return out, c.value, r.value


def ref_mixed_frame(srcs, width, height, level, mode, palette="standard", pad=False):
    """the reference's own create_mixed_ascii_frame_for_client (src/server/stream.c:958-1191), run end to end
    inside oracle/_ref (oracle/ref_stream_shim.c feeds it synthetic clients).  srcs[i] may be None (client
    connected, no video).  Returns (bytes | None, out_size, sources_with_video)."""
    k = len(srcs)
    arrs = [None if s is None else np.ascontiguousarray(s, np.uint8) for s in srcs]
    u8p = C.POINTER(C.c_uint8)
    ptrs = (u8p * k)(*[u8p() if a is None else a.ctypes.data_as(u8p) for a in arrs])
    ws = (C.c_int * k)(*[0 if a is None else a.shape[1] for a in arrs])
    hs = (C.c_int * k)(*[0 if a is None else a.shape[0] for a in arrs])
    caps = make_caps(level, mode, pad)
    n, cnt = C.c_size_t(0), C.c_int(0)
    f = ref().ref_oracle_mixed_frame
    f.restype = C.c_void_p
    r = f(ptrs, ws, hs, k, width, height, C.byref(caps), pal_bytes(palette), C.byref(n), C.byref(cnt))
    if not r:
        return None, n.value, cnt.value
    s = C.string_at(r, n.value)
    _libc.free(C.c_void_p(r))
    return s, n.value, cnt.value


def mixed_frame_fixup(s):
    """stream.c:1085-1127: a frame that does not end in ESC[0m is cut after its last ESC[0m (if it has one)"""
    rst = b"\x1b[0m"
    if len(s) >= 4 and not s.endswith(rst):
        k = s.rfind(rst)
        if k >= 0:
            return s[:k + 4]
    return s


def port_mixed_frame(srcs, width, height, level, mode, palette="standard", pad=False, scale=0):
    """the same entry restated on the port: count sources -> (the source | the W x 2H composite) ->
    orc_convert_caps(aspect=1, stretch=0) with h doubled for half-block (stream.c:829) -> reset fix-up"""
    live = [s for s in srcs if s is not None]
    if not live:
        return None, 0, 0
    comp = live[0] if len(live) == 1 else port_composite(live, width, height)[0]
    h = height * 2 if mode == 2 else height
    s = port_convert(comp, width, h, level, mode, palette, aspect=True, stretch=False, pad=pad, scale=scale)
    if s is None:
        return None, 0, len(live)
    s = mixed_frame_fixup(s)
    return s, len(s), len(live)


def composite_degenerate(srcs, width, height):
    """True when some source's fitted target is 0 px wide or tall.  The reference dereferences the NULL that
    image_new_from_pool(0, h) returns there (stream.c:723-749) and crashes, so such inputs have no reference
    answer; the product and the port skip that source instead (DESIGN.md §2 divergences)."""
    live = [s for s in srcs if s is not None]
    if len(live) < 2:
        return False
    k = len(live)  # the layout is computed for ALL sources with video (stream.c:670), only the first 9 are placed (:687)
    ws = (C.c_int * k)(*[s.shape[1] for s in live])
    hs = (C.c_int * k)(*[s.shape[0] for s in live])
    c, r = C.c_int(0), C.c_int(0)
    port().orc_grid_layout(ws, hs, k, width, height, C.byref(c), C.byref(r))
    live = live[:9]
    cw, ch = width // c.value, (height * 2) // r.value
    if cw <= 0 or ch <= 0:
        return True
    f32 = np.float32
    for s in live:
        a = f32(s.shape[1]) / f32(s.shape[0])
        if a > f32(cw) / f32(ch):
            tw, th = cw, int(f32(cw) / a + f32(0.5))
        else:
            tw, th = int(f32(ch) * a + f32(0.5)), ch
        if tw <= 0 or th <= 0:
            return True
    return False


def mixed_cases(count=60, seed=11):
    """deterministic server-path cases: per client (pattern, w, h) or None, terminal W x H, caps"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        n = int(rng.integers(1, 12))
        clients = []
        for i in range(n):
            if rng.random() < 0.2:
                clients.append(None)
            else:
                clients.append([("noise", "bars", "gradient", "grey")[int(rng.integers(0, 4))],
                                int(rng.integers(32, 480)), int(rng.integers(24, 360))])
        case = dict(clients=clients, W=int(rng.integers(30, 220)), H=int(rng.integers(10, 70)),
                    level=int(rng.integers(0, 4)), mode=int(rng.integers(0, 3)), pad=int(rng.integers(0, 2)),
                    palette=("standard", "blocks", "cool")[int(rng.integers(0, 3))])
        if composite_degenerate(mixed_sources(case), case["W"], case["H"]):
            continue
        out.append(case)
    # hand-picked: nobody sends video; one 1080p sender; nine 720p senders on a large terminal
    out.append(dict(clients=[None, None], W=80, H=24, level=3, mode=0, pad=1, palette="standard"))
    out.append(dict(clients=[None, ["noise", 1920, 1080]], W=203, H=61, level=3, mode=2, pad=1, palette="standard"))
    out.append(dict(clients=[["noise", 1280, 720]] * 9, W=240, H=67, level=2, mode=0, pad=0, palette="standard"))
    out.append(dict(clients=[["gradient", 640, 480]] * 4, W=160, H=48, level=1, mode=1, pad=1, palette="blocks"))
    return out


def mixed_sources(case):
    return [None if c is None else gen(c[0], c[1], c[2], i) for i, c in enumerate(case["clients"])]


# ----------------------------------------------------------------------------- client display path (8f rows 1, 3)
FILTERS = {"none": 0, "black": 1, "white": 2, "green": 3, "magenta": 4, "fuchsia": 5, "orange": 6, "teal": 7,
           "cyan": 8, "pink": 9, "red": 10, "yellow": 11, "rainbow": 12}


def ref_display_convert(img, cols, rows, level, mode, palette="standard", aspect=False, stretch=False, pad=False,
                        flip_x=False, flip_y=False, color_filter=0, time_s=0.0):
    """flip -> apply_color_filter -> ascii_convert_with_capabilities -> rainbow_replace_ansi_colors, all reference
    code, in the order of session_display_convert_to_ascii (display.c:484-671); see oracle/ref_display_shim.c"""
    a, p = as_u8(img)
    caps = make_caps(level, mode, pad)
    n = C.c_size_t(0)
    return _take(ref().ref_oracle_display_convert(p, a.shape[1], a.shape[0], cols, rows, C.byref(caps), int(aspect),
                                                  int(stretch), pal_bytes(palette), int(flip_x), int(flip_y),
                                                  int(color_filter), float(time_s), C.byref(n)))


def port_display_convert(img, cols, rows, level, mode, palette="standard", aspect=False, stretch=False, pad=False,
                         flip_x=False, flip_y=False, color_filter=0, time_s=0.0, scale=SCALE_NN):
    a, p = as_u8(img)
    n = C.c_size_t(0)
    return _take(port().orc_display_convert(p, a.shape[1], a.shape[0], cols, rows, level, mode, int(pad), int(aspect),
                                            int(stretch), pal_bytes(palette), int(flip_x), int(flip_y),
                                            int(color_filter), float(time_s), scale, C.byref(n)))


def ref_color_filter(img, color_filter, time_s=0.0):
    a = np.array(img, dtype=np.uint8, copy=True, order="C")
    rc = ref().apply_color_filter(a.ctypes.data_as(C.POINTER(C.c_uint8)), a.shape[1], a.shape[0], a.shape[1] * 3,
                                  int(color_filter), float(time_s))
    return rc, a


def port_color_filter(img, color_filter, time_s=0.0):
    a = np.array(img, dtype=np.uint8, copy=True, order="C")
    rc = port().orc_apply_color_filter(a.ctypes.data_as(C.POINTER(C.c_uint8)), a.shape[1], a.shape[0],
                                       a.shape[1] * 3, int(color_filter), float(time_s))
    return rc, a


def rainbow_rgb(lib_fn, t):
    r, g, b = C.c_uint8(0), C.c_uint8(0), C.c_uint8(0)
    lib_fn(float(t), C.byref(r), C.byref(g), C.byref(b))
    return r.value, g.value, b.value


def display_cases(count=48, seed=23):
    """deterministic client-display cases: geometry x caps x flips x filter x time"""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        out.append(dict(pattern=("noise", "bars", "gradient", "grey")[int(rng.integers(0, 4))],
                        W=int(rng.integers(24, 400)), H=int(rng.integers(16, 300)),
                        cols=int(rng.integers(8, 120)), rows=int(rng.integers(4, 50)),
                        level=int(rng.integers(0, 4)), mode=int(rng.integers(0, 3)),
                        aspect=int(rng.integers(0, 2)), pad=int(rng.integers(0, 2)),
                        flip_x=int(rng.integers(0, 2)), flip_y=int(rng.integers(0, 2)),
                        filter=int(rng.integers(0, 13)), time=round(float(rng.random() * 20.0), 3),
                        palette=("standard", "blocks", "cool")[int(rng.integers(0, 3))]))
    # hand-picked: BASELINE shapes with the filters that matter, rainbow on both truecolor grammars, 1-px-wide source
    out.append(dict(pattern="noise", W=3840, H=2160, cols=320, rows=96, level=3, mode=2, aspect=0, pad=0, flip_x=1,
                    flip_y=0, filter=3, time=1.25, palette="standard"))
    out.append(dict(pattern="noise", W=3840, H=2160, cols=320, rows=96, level=3, mode=2, aspect=1, pad=1, flip_x=0,
                    flip_y=1, filter=12, time=2.5, palette="standard"))
    out.append(dict(pattern="noise", W=1920, H=1080, cols=160, rows=48, level=3, mode=0, aspect=0, pad=0, flip_x=1,
                    flip_y=1, filter=12, time=0.4, palette="standard"))
    out.append(dict(pattern="bars", W=1920, H=1080, cols=160, rows=48, level=2, mode=0, aspect=0, pad=0, flip_x=0,
                    flip_y=0, filter=1, time=0.0, palette="blocks"))
    out.append(dict(pattern="noise", W=1, H=9, cols=4, rows=3, level=3, mode=0, aspect=0, pad=0, flip_x=1,
                    flip_y=1, filter=5, time=0.0, palette="standard"))
    out.append(dict(pattern="gradient", W=640, H=480, cols=80, rows=24, level=0, mode=0, aspect=0, pad=0, flip_x=1,
                    flip_y=0, filter=12, time=3.3, palette="standard"))
    return out


def display_args(case):
    return dict(cols=case["cols"], rows=case["rows"], level=case["level"], mode=case["mode"], palette=case["palette"],
                aspect=bool(case["aspect"]), stretch=False, pad=bool(case["pad"]), flip_x=bool(case["flip_x"]),
                flip_y=bool(case["flip_y"]), color_filter=case["filter"], time_s=case["time"])


# ----------------------------------------------------------------------------- wire packaging (8f row 4)
def ref_packet_header(frame, width, height):
    out = (C.c_uint8 * 24)()
    ref().ref_oracle_frame_packet_header(frame, len(frame), width, height, out)
    return bytes(out)


def port_packet_header(frame, width, height):
    out = (C.c_uint8 * 24)()
    port().orc_frame_packet_header(frame, len(frame), width, height, out)
    return bytes(out)
