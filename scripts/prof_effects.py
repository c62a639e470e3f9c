"""Small fixed workload for ncu: the kernels of effects.cu — whole-image colour filter on 16 resident 4K frames,
CRC32-C + packet headers over a 32-frame 4K half-block batch.  Never used for reported numbers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

assert acb.lib().acb200_init(0) == 0
k = 16
img = torch.randint(0, 256, (k * 2160, 3840, 3), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(4):
    assert acb.color_filter_device(img.data_ptr(), 3840, k * 2160, 3840 * 3, 3) == 0
acb.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = acb.make_cfg(3840, 2160, 320, 192, 3, 2, "standard", scale=acb.SCALE_BOX)
cap = acb.frame_capacity(cfg)
d_in = torch.randint(0, 256, (n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
d_len = torch.empty(n, dtype=torch.int32, device="cuda")
d_hdr = torch.empty(n * 24, dtype=torch.uint8, device="cuda")
d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
acb.render_batch_device(cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
for _ in range(4):
    acb.frame_packets_device(d_out.data_ptr(), cap, d_len.data_ptr(), n, 320, 96, d_hdr.data_ptr())
acb.synchronize()
print("ok", int(d_len.sum().item()), "string bytes")
