"""why does a low-entropy batch have a longer pass than its kernel time? per-iteration totals for several iteration counts"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import ascii_chat_b200 as acb
assert acb.lib().acb200_init(0) == 0
n, W, H = 256, 3840, 2160
for content in ("flat", "noise", "flat"):
    if content == "noise":
        d_in = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    else:
        band = torch.randint(0, 256, (n, H // 40 + 1, 1, 3), dtype=torch.uint8, device="cuda")
        d_in = band.repeat_interleave(40, dim=1)[:, :H].expand(n, H, W, 3).contiguous()
    cfg = acb.make_cfg(W, H, 320, 192, 3, 2, "standard", scale=acb.SCALE_BOX)
    cap = acb.frame_capacity(cfg)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    a = (cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
    acb.time_batch_device(*a, 3)
    for iters in (1, 2, 10, 40):
        tot, ker = acb.time_batch_device(*a, iters)
        print(content, "iters", iters, "tot/iter %.4f ms" % (tot / iters), "ker/iter %.4f ms" % (ker / iters))
    del d_in
