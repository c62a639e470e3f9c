"""Resident-batch throughput of the BASELINE configs in both downscale modes (device-timed); writes one JSON object.
Not the headline bench: context numbers for DESIGN.md / profiles."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import ascii_chat_b200 as acb
assert acb.lib().acb200_init(0) == 0
out = []
CASES = [("C1 640x480->80x24 mono fg", 640, 480, 80, 24, 0, 0, 2048),
         ("C2 1920x1080->160x48 ANSI-256 fg", 1920, 1080, 160, 48, 2, 0, 1024),
         ("C3 3840x2160->320x96 truecolor half-block", 3840, 2160, 320, 96, 3, 2, 256),
         ("3840x2160->320x96 truecolor fg", 3840, 2160, 320, 96, 3, 0, 256)]
ONLY = os.environ.get("MEASURE_ONLY")      # substring of the config name, e.g. "320x96" (tuning aid)
SCALES = os.environ.get("MEASURE_SCALES", "box,nn").split(",")
for name, W, H, c, r, level, mode, n in CASES:
    if ONLY and ONLY not in name:
        continue
    for content in ("noise", "flat"):
        if content == "noise":
            d_in = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
        else:  # low-entropy frames: horizontal colour bands, long runs (typical of real video vs. the noise worst case)
            band = torch.randint(0, 256, (n, H // 40 + 1, 1, 3), dtype=torch.uint8, device="cuda")
            d_in = band.repeat_interleave(40, dim=1)[:, :H].expand(n, H, W, 3).contiguous()
        for scale, sname in ((acb.SCALE_BOX, "box"), (acb.SCALE_NN, "nn")):
            if sname not in SCALES:
                continue
            cfg = acb.make_cfg(W, H, c, r * 2 if mode == 2 else r, level, mode, "standard", scale=scale)
            cap = acb.frame_capacity(cfg)
            d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
            d_len = torch.empty(n, dtype=torch.int32, device="cuda")
            d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
            a = (cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
            acb.time_batch_device(*a, 12)  # fresh allocations: let first-touch effects settle
            tot, ker = acb.time_batch_device(*a, 10)
            ms = tot / 10
            rec = dict(config=name, content=content, downscale=sname, frames=n, ms_per_pass=ms, frames_per_s=n / ms * 1e3,
                       mpix_s=n * W * H / 1e6 / ms * 1e3, gbs_alg_3Bpx=n * W * H * 3 / (ker / 10 * 1e-3) / 1e9,
                       out_bytes_per_frame=float(d_len.float().mean().item()))
            out.append(rec)
            print(rec)
        del d_in
        torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
