"""What the box's PCIe link delivers to a plain pinned cudaMemcpy — the ceiling the e2e call's string traffic runs into.
Prints D2H and H2D GB/s for 1.2 MB (one frame's string), 16 MB and 256 MB transfers, alone and with 8 streams in flight."""
import torch

assert torch.cuda.is_available()
dev = torch.device("cuda", 0)
for nbytes in (1180548, 16 << 20, 256 << 20):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, dst, src in (("D2H", h, d), ("H2D", d, h)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(4, (1 << 30) // nbytes)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        print("%s %9d B x %4d, one stream : %6.1f GB/s" % (name, nbytes, reps, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9))
# 8 concurrent streams of frame-sized D2H copies (what 8+ caller threads generate)
nbytes, S = 1180548, 8
ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(S)]
hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(S)]
st = [torch.cuda.Stream() for _ in range(S)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in st:
    s.wait_stream(torch.cuda.current_stream())
reps = 400
for r in range(reps):
    for k in range(S):
        with torch.cuda.stream(st[k]):
            hs[k].copy_(ds[k], non_blocking=True)
for s in st:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
print("D2H %9d B x %4d x %d streams   : %6.1f GB/s" % (nbytes, reps, S, nbytes * reps * S / (e0.elapsed_time(e1) * 1e-3) / 1e9))
