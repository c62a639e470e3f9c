"""CRC32-C + packet header on small batches (the one-frame server path): row form vs segment form, event-timed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import ascii_chat_b200 as acb  # noqa: E402

assert acb.lib().acb200_init(0) == 0
ts = torch.cuda.Stream()
st = ts.cuda_stream
pitch = 1259904
for n, L in ((1, 1180548), (1, 90000), (1, 2000), (8, 1180548), (32, 1180548)):
    d_out = torch.randint(0, 256, (n * pitch,), dtype=torch.uint8, device="cuda")
    d_len = torch.full((n,), L, dtype=torch.int32, device="cuda")
    d_hdr = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        for _ in range(20):
            acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), n, 320, 96, d_hdr.data_ptr(), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts.synchronize()
        e0.record(ts)
        for _ in range(200):
            acb.frame_packets_device(d_out.data_ptr(), pitch, d_len.data_ptr(), n, 320, 96, d_hdr.data_ptr(), st)
        e1.record(ts)
        ts.synchronize()
    print("%s: %2d frame(s) of %7d bytes: %.2f us per call" % (os.environ.get("ACB200_CRC_KERNEL", "rows"), n, L,
                                                                 1e3 * e0.elapsed_time(e1) / 200), flush=True)
