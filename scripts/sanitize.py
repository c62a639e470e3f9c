"""Small ladder for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep shapes tiny)."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import ascii_chat_b200 as acb
import oracle_bind as ob
assert acb.lib().acb200_init(0) == 0
bad = 0
img = ob.gen("noise", 64, 48, 1); img[:, 20:30] = 0
for level, mode in itertools.product((0, 1, 2, 3), (0, 1, 2)):
    got = acb.ascii_convert_with_capabilities(img, 16, 8, acb.make_caps(level, mode), False, False, "blocks")
    bad += got != ob.port_convert(img, 16, 8, level, mode, "blocks")
img = ob.gen("bars", 320, 96, 2)
for level, mode in ((0, 0), (2, 0), (3, 0), (3, 2), (1, 2)):
    cfg = acb.make_cfg(320, 96, 40, 24 if mode == 2 else 12, level, mode, "standard", scale=acb.SCALE_BOX, pad_left=2, pad_top=1)
    got = acb.render_batch_host(cfg, [img, img[::-1].copy(), img])
    exp = ob.port_convert(img, 40, 12, level, mode, "standard", scale=ob.SCALE_BOX)
    exp = ob._take(ob.port().orc_pad_height(ob._take(ob.port().orc_pad_width(exp, 2)), 1))
    bad += got[0] != exp or got[2] != exp
srcs = [ob.port_convert(ob.gen("noise", 96, 64, i), 20, 6, 3, 0) for i in range(4)]
bad += acb.ascii_create_grid(srcs, 60, 20) != ob.port_create_grid(srcs, 60, 20)
got, _, _ = acb.composite([ob.gen("bars", 80, 60, i) for i in range(3)], 60, 20)
bad += not np.array_equal(got, ob.port_composite([ob.gen("bars", 80, 60, i) for i in range(3)], 60, 20)[0])
# display path (flip / filter / rainbow), whole-image filter, wire packaging, trailing-reset cut
img = ob.gen("noise", 64, 48, 3)
for level, mode, filt, fx, fy in ((3, 2, 3, 1, 0), (3, 0, 12, 0, 1), (2, 0, 1, 1, 1), (0, 0, 9, 1, 0)):
    got = acb.display_convert(img, 16, 8, acb.make_caps(level, mode), False, False, "standard", bool(fx), bool(fy), filt, 0.7)
    bad += got != ob.port_display_convert(img, 16, 8, level, mode, flip_x=fx, flip_y=fy, color_filter=filt, time_s=0.7)
    cfg = acb.make_cfg(64, 48, 16, 16 if mode == 2 else 8, level, mode, scale=acb.SCALE_BOX, flip_x=fx, flip_y=fy,
                       color_filter=filt, filter_time=0.7)
    bad += acb.render_batch_host(cfg, [img])[0] != ob.port_display_convert(
        img, 16, 8, level, mode, flip_x=fx, flip_y=fy, color_filter=filt, time_s=0.7, scale=ob.SCALE_BOX)
for (w, h) in ((37, 5), (64, 48)):
    src = ob.gen("noise", w, h, 9)
    for f in (1, 5, 12):
        bad += not np.array_equal(acb.apply_color_filter(src, f, 0.3)[1], ob.port_color_filter(src, f, 0.3)[1])
srcs = [ob.gen(("noise", "bars", "gradient")[i], 96, 64, i) for i in range(3)]
for i, s_ in enumerate(srcs):
    acb.source_update(i, s_)
for level, mode in ((3, 2), (0, 0), (2, 0)):
    caps = acb.make_caps(level, mode, True)
    exp = ob.port_mixed_frame(srcs, 40, 12, level, mode, "standard", True)
    bad += acb.mixed_frame([0, 1, 2], 40, 12, caps, "standard") != exp
    pkt = acb.mixed_frame_packet([0, 1, 2], 40, 12, caps, "standard")
    bad += pkt[0] != ob.port_packet_header(exp[0], 40, 12) + exp[0]
# round 2: the foreground-only dithered printers, rainbow replace on a string, the in-process grid, pinned ingest,
# pixel-space composition from pre-fitted cell images
img = ob.gen("noise", 37, 11, 5)
for variant in (0, 1, 2):
    bad += acb.image_print_16color_dithered(img, "standard", None if variant == 2 else variant == 0) != \
        ob.port_print_dither(img, "standard", variant)
s_true = ob.port_convert(ob.gen("noise", 96, 64, 2), 40, 12, 3, 0)
bad += acb.rainbow_replace_ansi_colors(s_true * 3, 1.1) != ob.port_rainbow_replace(s_true * 3, 1.1)
bad += acb.rainbow_replace_ansi_colors(b"no colour", 1.1) is not None
for i, s_ in enumerate(srcs):
    acb.source_update_pinned(i, s_)
cells = [ob.port_convert(s_, 20, 6, 2, 0) + b"\0" for s_ in srcs]
bad += acb.grid_frame([0, 1, 2], 20, 6, acb.make_caps(2, 0), "standard", 60, 20) != ob.port_create_grid(cells, 60, 20)
import torch
ws, hs = [s_.shape[1] for s_ in srcs], [s_.shape[0] for s_ in srcs]
d_src = [torch.from_numpy(s_).cuda() for s_ in srcs]
d_cell = []
for i in range(3):
    tw, th = acb.mixed_cell_size(ws, hs, i, 40, 12)
    d_cell.append(torch.zeros(max(1, tw * th * 3), dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    acb.resize_nn_device(d_src[i].data_ptr(), ws[i], hs[i], d_cell[i].data_ptr(), tw, th, None)
acb.synchronize()
got = acb.mixed_frame_device([t.data_ptr() for t in d_cell], ws, hs, True, 40, 12, acb.make_caps(2, 0, True), "standard")
bad += got != ob.port_mixed_frame(srcs, 40, 12, 2, 0, "standard", True)[0]
for cols, rows, filt, frames in ob.rain_sequences()[3:]:
    a, b = acb.DigitalRain(cols, rows, filt), ob.PortRain(cols, rows, filt)
    for s_, dt in frames[:4]:
        bad += a.apply(s_, dt) != b.apply(s_, dt)
    a.close()
    b.close()
# page-locked frames fetched by the device (k_gather_nn_rows on mapped host memory), unaligned frame starts, flips
buf = np.zeros(97 * 61 * 3 + 256, np.uint8)
assert acb.lib().acb200_register_host_memory(buf.ctypes.data, buf.nbytes) == 0
acb.lib().acb200_set_fetch_depth(-1)
for off, fx, fy in ((0, 0, 0), (1, 1, 0), (7, 0, 1), (16, 1, 1)):
    im = buf[off:off + 97 * 61 * 3].reshape(61, 97, 3)
    im[...] = ob.gen("noise", 97, 61, off)
    got = acb.display_convert(im, 40, 12, acb.make_caps(3, 2), False, False, "standard", bool(fx), bool(fy), 3, 0.7)
    bad += got != ob.port_display_convert(im, 40, 12, 3, 2, flip_x=fx, flip_y=fy, color_filter=3, time_s=0.7)
acb.lib().acb200_set_fetch_depth(0)
assert acb.lib().acb200_unregister_host_memory(buf.ctypes.data) == 0
lens = [0, 1, 63, 64, 65, 511, 512, 513, 2048 + 17, 16384, 16385, 40000]
pitch = 40016
arena = np.random.default_rng(1).integers(0, 256, (len(lens), pitch), dtype=np.uint8)
d_out, d_len = torch.from_numpy(arena).cuda(), torch.tensor(lens, dtype=torch.int32, device="cuda")
d_hdr = torch.zeros(len(lens) * 24, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for form in (2, 1):  # segment form, then the row form (the ladder's arena is small: the default would pick segments)
    acb.lib().acb200_set_crc_form(form)
    acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), len(lens), 80, 24, d_hdr.data_ptr(), None)
    acb.synchronize()
    hdr = d_hdr.cpu().numpy().reshape(len(lens), 24)
    for i, L in enumerate(lens):
        bad += hdr[i].tobytes() != ob.port_packet_header(arena[i, :L].tobytes(), 80, 24)
acb.lib().acb200_set_crc_form(0)
acb.trailing_reset_fixup_device(d_out.data_ptr(), pitch, d_len.data_ptr(), len(lens), None)
acb.synchronize()
hdr = d_hdr.cpu().numpy().reshape(len(lens), 24)
for i, L in enumerate(lens):
    bad += hdr[i].tobytes() != ob.port_packet_header(arena[i, :L].tobytes(), 80, 24)
print("sanitize ladder mismatches:", bad)
