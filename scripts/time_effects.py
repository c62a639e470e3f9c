"""Event-timed effects kernels on resident data (tuning aid): colour filter on 64 4K frames, CRC/packets on a 256-frame
4K half-block batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

assert acb.lib().acb200_init(0) == 0
ts = torch.cuda.Stream()
st = ts.cuda_stream


def timed(fn, iters=10, warm=3):
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts.synchronize()
        e0.record(ts)
        for _ in range(iters):
            fn()
        e1.record(ts)
        ts.synchronize()
    return e0.elapsed_time(e1) / iters


k = 64
img = torch.randint(0, 256, (k * 2160, 3840, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: acb.color_filter_device(img.data_ptr(), 3840, k * 2160, 3840 * 3, 3, 0.0, st))
print("color_filter: %.4f ms  %.0f GB/s (read+write)" % (ms, 2 * img.numel() / (ms * 1e-3) / 1e9))
del img
n = 256
cfg = acb.make_cfg(3840, 2160, 320, 192, 3, 2, "standard", scale=acb.SCALE_BOX)
cap = acb.frame_capacity(cfg)
d_in = torch.randint(0, 256, (n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
d_len = torch.empty(n, dtype=torch.int32, device="cuda")
d_hdr = torch.empty(n * 24, dtype=torch.uint8, device="cuda")
d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
acb.render_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr(), st)
ts.synchronize()
ms = timed(lambda: acb.frame_packets_device(d_out.data_ptr(), cap, d_len.data_ptr(), n, 320, 96, d_hdr.data_ptr(), st))
sb = int(d_len.sum().item())
print("frame_packets: %.4f ms  %.0f GB/s of string bytes" % (ms, sb / (ms * 1e-3) / 1e9))
# filtered box render (generic kernel today) vs plain box render on the same 64 frames
n = 64
for filt in (0, 3):
    cfg = acb.make_cfg(3840, 2160, 320, 192, 3, 2, "standard", scale=acb.SCALE_BOX, color_filter=filt)
    cap = acb.frame_capacity(cfg)
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    tot, ker = acb.time_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr(), 3)
    print("box render, colour filter %d: %.3f ms per %d frames  %.0f GB/s of source bytes" % (
        filt, tot / 3, n, n * 3840 * 2160 * 3 / (tot / 3 * 1e-3) / 1e9))
