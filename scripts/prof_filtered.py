"""ncu target: the 4K -> 320x96 truecolor half-block box render with a colour filter fused in (k_render_rows_ws2<4, true>),
64 resident frames, a few passes.  Prints the device-timed bandwidth (never a reported number when run under ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
assert acb.lib().acb200_init(0) == 0
for filt in (0, 3, 1):
    cfg = acb.make_cfg(3840, 2160, 320, 192, 3, 2, "standard", scale=acb.SCALE_BOX, color_filter=filt)
    cap = acb.frame_capacity(cfg)
    d_in = torch.randint(0, 256, (n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    a = (cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
    acb.time_batch_device(*a, 5)
    tot, ker = acb.time_batch_device(*a, 10)
    print("filter %d: %.3f ms per %d frames, %.0f GB/s of source bytes" % (filt, ker / 10, n, n * 3840 * 2160 * 3 / (ker / 10 * 1e-3) / 1e9))
