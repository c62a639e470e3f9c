"""ctypes mirror of include/asciichat_b200.h.

Function names, argument meaning and error behaviour are the reference's (libasciichat):
``ascii_convert_with_capabilities`` returns the frame string or ``None`` where the C function
returns NULL.  Buffers are numpy (host) or raw device pointers (batch API); torch is used by
callers only for device memory and process groups.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", os.environ.get("ACB200_LIB_NAME", "libasciichat_b200.so"))  # experiments: other name

TERM_COLOR_NONE, TERM_COLOR_16, TERM_COLOR_256, TERM_COLOR_TRUECOLOR = 0, 1, 2, 3
RENDER_MODE_FOREGROUND, RENDER_MODE_BACKGROUND, RENDER_MODE_HALF_BLOCK = 0, 1, 2
SCALE_NN, SCALE_BOX = 0, 1

PALETTE_CHARS_STANDARD = "   ...',;:clodxkO0KXNWM"  # palette.h:161
PALETTE_CHARS_BLOCKS = "   ░░▒▒▓▓██"
PALETTE_CHARS_DIGITAL = "   -=≡≣▰▱◼"
PALETTE_CHARS_MINIMAL = "   .-+*#"
PALETTE_CHARS_COOL = "   ▁▂▃▄▅▆▇█"
PALETTES = {"standard": PALETTE_CHARS_STANDARD, "blocks": PALETTE_CHARS_BLOCKS, "digital": PALETTE_CHARS_DIGITAL,
            "minimal": PALETTE_CHARS_MINIMAL, "cool": PALETTE_CHARS_COOL}


class image_t(C.Structure):  # image.h:143-148
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("pixels", C.c_void_p), ("alloc_method", C.c_uint8)]


class terminal_capabilities_t(C.Structure):  # platform/terminal.h:707-738
    _fields_ = [
        ("color_level", C.c_int), ("capabilities", C.c_uint32), ("color_count", C.c_uint32),
        ("utf8_support", C.c_bool), ("detection_reliable", C.c_bool), ("render_mode", C.c_int),
        ("term_type", C.c_char * 64), ("colorterm", C.c_char * 64), ("wants_background", C.c_bool),
        ("palette_type", C.c_int), ("palette_custom", C.c_char * 64), ("desired_fps", C.c_uint8),
        ("color_filter", C.c_int), ("wants_padding", C.c_bool), ("pad_height", C.c_size_t),
    ]


class digital_rain_column_t(C.Structure):  # video/anim/digital_rain.h
    _fields_ = [("time_offset", C.c_float), ("speed_multiplier", C.c_float), ("phase_offset", C.c_float)]


class digital_rain_t(C.Structure):  # video/anim/digital_rain.h (previous_brightness: device memory here)
    _fields_ = [("columns", C.POINTER(digital_rain_column_t)), ("num_columns", C.c_int), ("num_rows", C.c_int),
                ("time", C.c_float), ("fall_speed", C.c_float), ("raindrop_length", C.c_float),
                ("brightness_decay", C.c_float), ("animation_speed", C.c_float), ("color_r", C.c_uint8),
                ("color_g", C.c_uint8), ("color_b", C.c_uint8), ("cursor_brightness", C.c_float),
                ("rainbow_mode", C.c_bool), ("first_frame", C.c_bool), ("previous_brightness", C.c_void_p)]


class ascii_frame_source_t(C.Structure):  # ascii.h:358-361
    _fields_ = [("frame_data", C.c_char_p), ("frame_size", C.c_size_t)]


class acb200_render_cfg_t(C.Structure):
    _fields_ = [("src_w", C.c_int), ("src_h", C.c_int), ("cols", C.c_int), ("rows_px", C.c_int),
                ("color_level", C.c_int), ("render_mode", C.c_int), ("scale", C.c_int), ("pad_left", C.c_int),
                ("pad_top", C.c_int), ("palette", C.c_char_p),
                ("flip_x", C.c_int), ("flip_y", C.c_int), ("color_filter", C.c_int), ("filter_time", C.c_float)]


EXPORTS = [  # every symbol include/asciichat_b200.h declares
    "ascii_convert", "ascii_convert_with_capabilities", "image_print_with_capabilities", "image_resize",
    "image_print", "image_print_color", "image_print_256color", "image_print_16color",
    "image_print_16color_dithered_with_background", "rgb_to_truecolor_halfblocks_scalar", "rgb_to_halfblocks_scalar",
    "rgb_to_16color_halfblocks_scalar", "rgb_to_256color_halfblocks_scalar", "ascii_create_grid", "ascii_simd_init",
    "simd_caches_destroy_all", "acb200_init", "acb200_shutdown", "acb200_last_error", "acb200_last_error_message",
    "acb200_set_allocator", "acb200_set_option_render_mode", "acb200_set_default_scale", "acb200_frame_capacity",
    "acb200_scratch_bytes", "acb200_render_batch_device", "acb200_render_batch_host", "acb200_time_batch_device",
    "acb200_composite_host", "acb200_grid_layout", "acb200_aspect_ratio", "acb200_launch_count", "acb200_version",
    "acb200_create_grid_device", "acb200_synchronize", "acb200_source_update", "acb200_source_clear",
    "acb200_mixed_frame", "apply_color_filter", "color_filter_calculate_rainbow", "acb200_display_convert",
    "acb200_color_filter_device", "acb200_frame_packets_device", "acb200_mixed_frame_packet",
    "acb200_trailing_reset_fixup_device", "acb200_source_update_wire",
    "acb200_init_devices", "acb200_device_count", "acb200_device_at", "acb200_bind_thread", "acb200_thread_device",
    "acb200_set_sync_mode", "acb200_source_acquire", "acb200_source_commit", "acb200_source_device",
    "acb200_grid_frame", "acb200_quantize_table_device", "image_print_16color_dithered",
    "rainbow_replace_ansi_colors", "acb200_mixed_cell_size", "acb200_resize_nn_device", "acb200_mixed_frame_device",
    "digital_rain_init", "digital_rain_destroy", "digital_rain_apply", "digital_rain_reset",
    "digital_rain_set_fall_speed", "digital_rain_set_raindrop_length", "digital_rain_set_color",
    "digital_rain_set_color_from_filter", "acb200_host_phase_stats",
    "acb200_register_host_memory", "acb200_unregister_host_memory", "acb200_set_fetch_depth",
    "acb200_nn_row_schedule", "acb200_set_crc_form",
]


def build_library(force=False):
    import importlib.util
    spec = importlib.util.spec_from_file_location("acb200_build", os.path.join(HERE, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build(force=force)


_lib = None
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def lib():
    """load libasciichat_b200.so (raises if it was never built: there is no fallback)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libasciichat_b200.so is not built: run `python ascii-chat_b200/build.py` "
                           "(or __graft_entry__.build()); this package has no non-CUDA path")
    L = C.CDLL(LIB_PATH)
    ip, cp = C.POINTER(image_t), C.POINTER(terminal_capabilities_t)
    u8p = C.POINTER(C.c_uint8)
    cfgp = C.POINTER(acb200_render_cfg_t)
    L.ascii_convert_with_capabilities.restype = C.c_void_p
    L.ascii_convert_with_capabilities.argtypes = [ip, C.c_ssize_t, C.c_ssize_t, cp, C.c_bool, C.c_bool, C.c_char_p]
    L.ascii_convert.restype = C.c_void_p
    L.ascii_convert.argtypes = [ip, C.c_ssize_t, C.c_ssize_t, C.c_bool, C.c_bool, C.c_bool, C.c_char_p, C.c_char_p]
    L.image_print_with_capabilities.restype = C.c_void_p
    L.image_print_with_capabilities.argtypes = [ip, cp, C.c_char_p]
    L.image_resize.restype = None
    L.image_resize.argtypes = [ip, ip]
    for name in ("image_print", "image_print_color", "image_print_256color", "image_print_16color"):
        getattr(L, name).restype = C.c_void_p
        getattr(L, name).argtypes = [ip, C.c_char_p]
    L.image_print_16color_dithered_with_background.restype = C.c_void_p
    L.image_print_16color_dithered_with_background.argtypes = [ip, C.c_bool, C.c_char_p]
    L.rgb_to_truecolor_halfblocks_scalar.restype = C.c_void_p
    L.rgb_to_truecolor_halfblocks_scalar.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
    for name in ("rgb_to_halfblocks_scalar", "rgb_to_16color_halfblocks_scalar", "rgb_to_256color_halfblocks_scalar"):
        getattr(L, name).restype = C.c_void_p
        getattr(L, name).argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.ascii_create_grid.restype = C.c_void_p
    L.ascii_create_grid.argtypes = [C.POINTER(ascii_frame_source_t), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    L.ascii_simd_init.restype = None
    L.simd_caches_destroy_all.restype = None
    L.acb200_init.argtypes = [C.c_int]
    L.acb200_last_error_message.restype = C.c_char_p
    L.acb200_set_option_render_mode.argtypes = [C.c_int]
    L.acb200_set_default_scale.argtypes = [C.c_int]
    L.acb200_frame_capacity.restype = C.c_size_t
    L.acb200_frame_capacity.argtypes = [cfgp]
    L.acb200_scratch_bytes.restype = C.c_size_t
    L.acb200_scratch_bytes.argtypes = [cfgp, C.c_int]
    L.acb200_render_batch_device.argtypes = [cfgp, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
    L.acb200_render_batch_host.argtypes = [cfgp, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_size_t)]
    L.acb200_time_batch_device.argtypes = [cfgp, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.acb200_composite_host.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.acb200_grid_layout.restype = None
    L.acb200_grid_layout.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.acb200_aspect_ratio.restype = None
    L.acb200_aspect_ratio.argtypes = [C.c_ssize_t] * 4 + [C.c_bool, C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t)]
    L.acb200_launch_count.restype = C.c_uint64
    L.acb200_version.restype = C.c_char_p
    L.acb200_create_grid_device.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
    L.acb200_source_update.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.acb200_source_clear.argtypes = [C.c_int]
    L.acb200_source_update_wire.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.acb200_mixed_frame.restype = C.c_void_p
    L.acb200_mixed_frame.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_ushort, C.c_ushort,
                                     C.POINTER(terminal_capabilities_t), C.c_char_p, C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_int)]
    L.apply_color_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float]
    L.color_filter_calculate_rainbow.restype = None
    L.color_filter_calculate_rainbow.argtypes = [C.c_float, u8p, u8p, u8p]
    L.acb200_display_convert.restype = C.c_void_p
    L.acb200_display_convert.argtypes = [ip, C.c_ssize_t, C.c_ssize_t, cp, C.c_bool, C.c_bool, C.c_char_p, C.c_bool,
                                         C.c_bool, C.c_int, C.c_float]
    L.acb200_color_filter_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float,
                                             C.c_void_p]
    L.acb200_frame_packets_device.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p]
    L.acb200_trailing_reset_fixup_device.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    L.acb200_mixed_frame_packet.restype = C.c_void_p
    L.acb200_mixed_frame_packet.argtypes = L.acb200_mixed_frame.argtypes
    L.image_print_16color_dithered.restype = C.c_void_p
    L.image_print_16color_dithered.argtypes = [ip, C.c_char_p]
    L.rainbow_replace_ansi_colors.restype = C.c_void_p
    L.rainbow_replace_ansi_colors.argtypes = [C.c_char_p, C.c_float]
    L.acb200_quantize_table_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.acb200_mixed_cell_size.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_ushort,
                                         C.c_ushort, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.acb200_resize_nn_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.acb200_mixed_frame_device.restype = C.c_void_p
    L.acb200_mixed_frame_device.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                            C.c_int, C.c_ushort, C.c_ushort, C.POINTER(terminal_capabilities_t),
                                            C.c_char_p, C.POINTER(C.c_size_t)]
    rp = C.POINTER(digital_rain_t)
    L.digital_rain_init.restype = rp
    L.digital_rain_init.argtypes = [C.c_int, C.c_int]
    L.digital_rain_destroy.restype = None
    L.digital_rain_destroy.argtypes = [rp]
    L.digital_rain_apply.restype = C.c_void_p
    L.digital_rain_apply.argtypes = [rp, C.c_char_p, C.c_float]
    L.digital_rain_reset.restype = None
    L.digital_rain_reset.argtypes = [rp]
    for name in ("digital_rain_set_fall_speed", "digital_rain_set_raindrop_length"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [rp, C.c_float]
    L.digital_rain_set_color.restype = None
    L.digital_rain_set_color.argtypes = [rp, C.c_uint8, C.c_uint8, C.c_uint8]
    L.digital_rain_set_color_from_filter.restype = None
    L.digital_rain_set_color_from_filter.argtypes = [rp, C.c_int]
    L.acb200_init_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.acb200_device_at.argtypes = [C.c_int]
    L.acb200_bind_thread.argtypes = [C.c_int]
    L.acb200_set_sync_mode.restype = None
    L.acb200_set_sync_mode.argtypes = [C.c_int, C.c_int]
    L.acb200_register_host_memory.argtypes = [C.c_void_p, C.c_size_t]
    L.acb200_unregister_host_memory.argtypes = [C.c_void_p]
    L.acb200_nn_row_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]
    L.acb200_set_crc_form.restype = None
    L.acb200_set_crc_form.argtypes = [C.c_int]
    L.acb200_set_fetch_depth.restype = None
    L.acb200_set_fetch_depth.argtypes = [C.c_int]
    L.acb200_host_phase_stats.restype = None
    L.acb200_host_phase_stats.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    L.acb200_source_acquire.restype = C.c_void_p
    L.acb200_source_acquire.argtypes = [C.c_int, C.c_size_t]
    L.acb200_source_commit.argtypes = [C.c_int, C.c_int, C.c_int]
    L.acb200_source_device.argtypes = [C.c_int]
    L.acb200_grid_frame.restype = C.c_void_p
    L.acb200_grid_frame.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(terminal_capabilities_t),
                                    C.c_bool, C.c_bool, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    _lib = L
    return L


def last_error():
    L = lib()
    return L.acb200_last_error(), (L.acb200_last_error_message() or b"").decode("utf-8", "replace")


def _take(ptr):
    if not ptr:
        return None
    s = C.string_at(ptr)
    _libc.free(ptr)
    return s


def _pal(p):
    if p is None:
        return None
    return PALETTES.get(p, p).encode() if isinstance(p, str) else p


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    return a, image_t(a.shape[1], a.shape[0], a.ctypes.data, 0)


def make_caps(color_level, render_mode, wants_padding=False):
    c = terminal_capabilities_t()
    c.color_level, c.render_mode, c.utf8_support, c.wants_padding = color_level, render_mode, True, wants_padding
    return c


def make_cfg(src_w, src_h, cols, rows_px, color_level, render_mode, palette="standard", scale=SCALE_NN, pad_left=0,
             pad_top=0, flip_x=False, flip_y=False, color_filter=0, filter_time=0.0):
    return acb200_render_cfg_t(src_w, src_h, cols, rows_px, color_level, render_mode, scale, pad_left, pad_top,
                               _pal(palette), int(flip_x), int(flip_y), int(color_filter), float(filter_time))


# ---- the reference's entry points -------------------------------------------------------------
def ascii_convert_with_capabilities(image, width, height, caps, use_aspect_ratio, stretch, palette_chars):
    a, im = _img(image)
    return _take(lib().ascii_convert_with_capabilities(C.byref(im), width, height, C.byref(caps), use_aspect_ratio,
                                                       stretch, _pal(palette_chars)))


def ascii_convert(image, width, height, color, aspect_ratio, stretch, palette_chars, luminance_palette=None):
    a, im = _img(image)
    lum = luminance_palette if luminance_palette is not None else bytes(range(1, 256)) + b"\x01"
    return _take(lib().ascii_convert(C.byref(im), width, height, color, aspect_ratio, stretch, _pal(palette_chars), lum))


def image_print_with_capabilities(image, caps, palette):
    a, im = _img(image)
    return _take(lib().image_print_with_capabilities(C.byref(im), C.byref(caps), _pal(palette)))


def image_resize(image, dw, dh):
    a, src = _img(image)
    out = np.zeros((dh, dw, 3), np.uint8)
    dst = image_t(dw, dh, out.ctypes.data, 0)
    lib().image_resize(C.byref(src), C.byref(dst))
    return out


def ascii_create_grid(frames, width, height):
    arr = (ascii_frame_source_t * len(frames))()
    for i, f in enumerate(frames):
        arr[i].frame_data, arr[i].frame_size = (f, len(f)) if f is not None else (None, 0)
    n = C.c_size_t(0)
    r = lib().ascii_create_grid(arr, len(frames), width, height, C.byref(n))
    return _take(r), n.value


def composite(srcs, width, height):
    """server pixel-space composite (stream.c:664-779): list of (h,w,3) arrays -> (2*height, width, 3)"""
    arrs = [np.ascontiguousarray(s, np.uint8) for s in srcs]
    k = len(arrs)
    ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
    ws = (C.c_int * k)(*[a.shape[1] for a in arrs])
    hs = (C.c_int * k)(*[a.shape[0] for a in arrs])
    out = np.empty((height * 2, width, 3), np.uint8)
    c, r = C.c_int(0), C.c_int(0)
    rc = lib().acb200_composite_host(ptrs, ws, hs, k, width, height, out.ctypes.data, C.byref(c), C.byref(r))
    if rc:
        raise RuntimeError("acb200_composite_host failed: %s" % (last_error(),))
    return out, c.value, r.value


# ---- the server's per-client frame with resident sources (src/server/stream.c:958-1191) --------
MAX_SOURCES = 32


def source_update(slot, image):
    a = np.ascontiguousarray(image, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    return lib().acb200_source_update(slot, a.ctypes.data, a.shape[1], a.shape[0])


def source_update_wire(slot, payload):
    """payload: bytes of an IMAGE_FRAME packet body, [w:be32][h:be32][RGB24] (src/server/protocol.c:737-889)"""
    return lib().acb200_source_update_wire(slot, payload, len(payload))


def source_clear(slot):
    return lib().acb200_source_clear(slot)


def source_update_pinned(slot, image):
    """the zero-staging-copy ingest: receive into the slot's pinned buffer, then commit"""
    a = np.ascontiguousarray(image, dtype=np.uint8)
    p = lib().acb200_source_acquire(slot, a.nbytes)
    if not p:
        return last_error()[0] or -1
    C.memmove(p, a.ctypes.data, a.nbytes)  # stands in for the transport's recv() into the buffer
    return lib().acb200_source_commit(slot, a.shape[1], a.shape[0])


def mixed_cell_size(ws, hs, i, width, height):
    n = len(ws)
    tw, th = C.c_int(0), C.c_int(0)
    rc = lib().acb200_mixed_cell_size((C.c_int * n)(*ws), (C.c_int * n)(*hs), n, i, width, height, C.byref(tw),
                                      C.byref(th))
    if rc:
        raise RuntimeError("acb200_mixed_cell_size failed: %s" % (last_error(),))
    return tw.value, th.value


def resize_nn_device(d_src, sw, sh, d_dst, dw, dh, stream=None):
    rc = lib().acb200_resize_nn_device(d_src, sw, sh, d_dst, dw, dh, stream)
    if rc:
        raise RuntimeError("acb200_resize_nn_device failed: %s" % (last_error(),))


def mixed_frame_device(d_ptrs, ws, hs, prefit, width, height, caps, palette_chars):
    """-> bytes | None; the frames (or pre-fitted cell images) are on the calling thread's device"""
    n = len(d_ptrs)
    sz = C.c_size_t(0)
    r = lib().acb200_mixed_frame_device((C.c_void_p * n)(*d_ptrs), (C.c_int * n)(*ws), (C.c_int * n)(*hs), n,
                                        1 if prefit else 0, width, height, C.byref(caps), _pal(palette_chars),
                                        C.byref(sz))
    if not r:
        return None
    s = C.string_at(r, sz.value)
    _libc.free(r)
    return s


def init_devices(devices=None):
    """one process, several GPUs (acb200_init_devices); None = every visible device"""
    if devices is None:
        return lib().acb200_init_devices(None, 0)
    arr = (C.c_int * len(devices))(*devices)
    return lib().acb200_init_devices(arr, len(devices))


def grid_frame(slots, cell_w, cell_h, caps, palette_chars, grid_w, grid_h, use_aspect_ratio=False, stretch=False):
    """-> (bytes | None, out_size): the discovery host's tick (host.c:664-717) with resident sources"""
    k = len(slots)
    arr = (C.c_int * max(k, 1))(*slots)
    n = C.c_size_t(0)
    r = lib().acb200_grid_frame(arr, k, cell_w, cell_h, C.byref(caps), use_aspect_ratio, stretch, _pal(palette_chars),
                                grid_w, grid_h, C.byref(n))
    if not r:
        return None, n.value
    s = C.string_at(r)
    _libc.free(r)
    return s, n.value


def mixed_frame(slots, width, height, caps, palette_chars):
    """-> (bytes | None, out_size, sources_with_video), like create_mixed_ascii_frame_for_client"""
    k = len(slots)
    arr = (C.c_int * max(k, 1))(*slots)
    n, cnt = C.c_size_t(0), C.c_int(0)
    r = lib().acb200_mixed_frame(arr, k, width, height, C.byref(caps) if caps is not None else None,
                                 _pal(palette_chars), C.byref(n), C.byref(cnt))
    if not r:
        return None, n.value, cnt.value
    s = C.string_at(r, n.value)
    _libc.free(r)
    return s, n.value, cnt.value


def mixed_frame_packet(slots, width, height, caps, palette_chars):
    """-> (header||frame bytes | None, out_size, sources_with_video): acb200_mixed_frame + acip_send_ascii_frame's
    24-byte ascii_frame_packet_t (lib/network/acip/server.c:188-236)"""
    k = len(slots)
    arr = (C.c_int * max(k, 1))(*slots)
    n, cnt = C.c_size_t(0), C.c_int(0)
    r = lib().acb200_mixed_frame_packet(arr, k, width, height, C.byref(caps) if caps is not None else None,
                                        _pal(palette_chars), C.byref(n), C.byref(cnt))
    if not r:
        return None, n.value, cnt.value
    s = C.string_at(r, n.value)
    _libc.free(r)
    return s, n.value, cnt.value


# ---- the client's display conversion (src/common/session/display.c:484-671) ---------------------
COLOR_FILTERS = {"none": 0, "black": 1, "white": 2, "green": 3, "magenta": 4, "fuchsia": 5, "orange": 6, "teal": 7,
                 "cyan": 8, "pink": 9, "red": 10, "yellow": 11, "rainbow": 12}  # platform/terminal.h:601-627


def display_convert(image, width, height, caps, preserve_aspect_ratio, stretch, palette_chars, flip_x=False,
                    flip_y=False, color_filter=0, time_seconds=0.0):
    a, im = _img(image)
    return _take(lib().acb200_display_convert(C.byref(im), width, height, C.byref(caps), preserve_aspect_ratio, stretch,
                                              _pal(palette_chars), flip_x, flip_y, int(color_filter),
                                              float(time_seconds)))


def apply_color_filter(image, color_filter, time_seconds=0.0, stride=None):
    """in place on a copy, like the reference on its copy (display.c:612-623): -> (rc, filtered array)"""
    a = np.array(image, dtype=np.uint8, copy=True, order="C")
    st = a.shape[1] * 3 if stride is None else stride
    rc = lib().apply_color_filter(a.ctypes.data, a.shape[1], a.shape[0], st, int(color_filter), float(time_seconds))
    return rc, a


def rainbow_replace_ansi_colors(ansi_string, time_seconds):
    """color_filter.c:348-408 on a finished string; None where the reference returns NULL"""
    return _take(lib().rainbow_replace_ansi_colors(ansi_string, float(time_seconds)))


def image_print_16color_dithered(image, palette, use_background=None):
    """use_background None: image_print_16color_dithered (foreground.c:650); else ..._with_background(img, flag) (:752)"""
    a, im = _img(image)
    if use_background is None:
        return _take(lib().image_print_16color_dithered(C.byref(im), _pal(palette)))
    return _take(lib().image_print_16color_dithered_with_background(C.byref(im), bool(use_background), _pal(palette)))


class DigitalRain:
    """digital_rain_t (lib/video/anim/digital_rain.c) — state on the GPU, the string work on the device"""

    def __init__(self, cols, rows, color_filter=0):
        self.p = lib().digital_rain_init(cols, rows)
        if not self.p:
            raise RuntimeError("digital_rain_init failed: %s" % (last_error(),))
        lib().digital_rain_set_color_from_filter(self.p, int(color_filter))

    def apply(self, frame, delta_time):
        return _take(lib().digital_rain_apply(self.p, frame, float(delta_time)))

    def reset(self):
        lib().digital_rain_reset(self.p)

    def close(self):
        if self.p:
            lib().digital_rain_destroy(self.p)
            self.p = None


def calculate_rainbow(t):
    r, g, b = C.c_uint8(0), C.c_uint8(0), C.c_uint8(0)
    lib().color_filter_calculate_rainbow(float(t), C.byref(r), C.byref(g), C.byref(b))
    return r.value, g.value, b.value


def color_filter_device(d_pixels, width, height, stride, color_filter, time_seconds=0.0, stream=None):
    return lib().acb200_color_filter_device(d_pixels, width, height, stride, int(color_filter), float(time_seconds),
                                            stream)


def frame_packets_device(d_out, out_pitch, d_out_len, n, width, height, d_headers, stream=None):
    rc = lib().acb200_frame_packets_device(d_out, out_pitch, d_out_len, n, width, height, d_headers, stream)
    if rc:
        raise RuntimeError("acb200_frame_packets_device failed: %s" % (last_error(),))


def trailing_reset_fixup_device(d_out, out_pitch, d_out_len, n, stream=None):
    rc = lib().acb200_trailing_reset_fixup_device(d_out, out_pitch, d_out_len, n, stream)
    if rc:
        raise RuntimeError("acb200_trailing_reset_fixup_device failed: %s" % (last_error(),))


def aspect_ratio(img_w, img_h, width, height, stretch=False):
    ow, oh = C.c_ssize_t(0), C.c_ssize_t(0)
    lib().acb200_aspect_ratio(img_w, img_h, width, height, stretch, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


# ---- batch interface --------------------------------------------------------------------------
def render_batch_host(cfg, frames):
    """frames: list of contiguous uint8 arrays (h,w,3) on the host -> list of bytes"""
    arrs = [np.ascontiguousarray(f, np.uint8) for f in frames]
    k = len(arrs)
    ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
    outs = (C.c_void_p * k)()
    lens = (C.c_size_t * k)()
    rc = lib().acb200_render_batch_host(C.byref(cfg), ptrs, k, outs, lens)
    if rc:
        raise RuntimeError("acb200_render_batch_host failed: %s" % (last_error(),))
    res = []
    for i in range(k):
        res.append(C.string_at(outs[i], lens[i]))
        _libc.free(outs[i])
    return res


def render_batch_host_ptrs(cfg, ptr_array, k, outs, lens):
    """zero-overhead variant for timing loops: caller owns the ctypes arrays; strings must be freed"""
    return lib().acb200_render_batch_host(C.byref(cfg), ptr_array, k, outs, lens)


def free_strings(outs, k):
    for i in range(k):
        if outs[i]:
            _libc.free(outs[i])
            outs[i] = None


def frame_capacity(cfg):
    return lib().acb200_frame_capacity(C.byref(cfg))


def scratch_bytes(cfg, n):
    return lib().acb200_scratch_bytes(C.byref(cfg), n)


def synchronize():
    """wait for the calling thread's internal stream (used when stream=None was passed to the device API)"""
    rc = lib().acb200_synchronize()
    if rc:
        raise RuntimeError("acb200_synchronize failed: %s" % (last_error(),))


def render_batch_device(cfg, d_frames, n, d_out, out_pitch, d_out_len, d_scratch, stream=None):
    rc = lib().acb200_render_batch_device(C.byref(cfg), d_frames, n, d_out, out_pitch, d_out_len, d_scratch, stream)
    if rc:
        raise RuntimeError("acb200_render_batch_device failed: %s" % (last_error(),))


def time_batch_device(cfg, d_frames, n, d_out, out_pitch, d_out_len, d_scratch, iters):
    tot, ker = C.c_float(0), C.c_float(0)
    rc = lib().acb200_time_batch_device(C.byref(cfg), d_frames, n, d_out, out_pitch, d_out_len, d_scratch, iters,
                                        C.byref(tot), C.byref(ker))
    if rc:
        raise RuntimeError("acb200_time_batch_device failed: %s" % (last_error(),))
    return tot.value, ker.value


def create_grid_device(d_ptrs, sizes, width, height, d_out, stream=None):
    k = len(d_ptrs)
    ptrs = (C.c_void_p * k)(*d_ptrs)
    sz = (C.c_size_t * k)(*sizes)
    n = C.c_size_t(0)
    rc = lib().acb200_create_grid_device(ptrs, sz, k, width, height, d_out, C.byref(n), stream)
    if rc:
        raise RuntimeError("acb200_create_grid_device failed: %s" % (last_error(),))
    return n.value


def launch_count():
    return int(lib().acb200_launch_count())
