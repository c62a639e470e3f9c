"""Small fixed workload for ncu: the 4K -> 320x96 truecolor half-block box render over a 32-frame resident
batch, 3 passes (+ the NN variant once).  Never used for reported numbers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
content = sys.argv[2] if len(sys.argv) > 2 else "noise"  # "flat": horizontal colour bands, long runs
assert acb.lib().acb200_init(0) == 0
for scale in (acb.SCALE_BOX, acb.SCALE_NN):
    cfg = acb.make_cfg(3840, 2160, 320, 192, 3, 2, "standard", scale=scale)
    cap = acb.frame_capacity(cfg)
    if content == "flat":
        band = torch.randint(0, 256, (n, 2160 // 40 + 1, 1, 3), dtype=torch.uint8, device="cuda")
        d_in = band.repeat_interleave(40, dim=1)[:, :2160].expand(n, 2160, 3840, 3).contiguous()
    else:
        d_in = torch.randint(0, 256, (n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        acb.render_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
    tot, ker = acb.time_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(),
                                     d_scr.data_ptr(), 3)
    print("scale", scale, "ms/pass", tot / 3, "kernel ms", ker / 3, "GB/s", n * 3840 * 2160 * 3 / (ker / 3 * 1e-3) / 1e9)
