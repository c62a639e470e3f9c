#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mpixels/s through the fused
downscale -> luminance -> quantise -> glyph kernel at 3840x2160, % of the HBM roofline, 1/2/4/8 GPUs,
with the reference's CPU path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

Workload (config.workload): BASELINE configs[2]/[4] — 3840x2160 RGB24 -> 320x96 truecolor half-block cells,
box-filter downscale (every source pixel read once: 3 B/px algorithmic), a ring of 256 distinct frames
resident in HBM (6.37 GB per pass >> 126 MB L2, so no L2 flush is needed between iterations).
A step = one pass of the render path over the 256-frame batch.  Frames are independent, so ranks take
disjoint rings with no data-path collective ("scaling": "weak").

One JSON line is printed by rank 0.  See DESIGN.md §6 for what each field means.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRC_W, SRC_H, COLS, ROWS = 3840, 2160, 320, 96
LEVEL, MODE = 3, 2  # truecolor, half-block
RING = 256
FRAME_BYTES = SRC_W * SRC_H * 3
MPIX = SRC_W * SRC_H / 1e6
METRIC = "Mpixels/s fused RGB->glyph render at 4K"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def traffic_from_profiles():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("render_rows_4k_hb_box_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_leg(threads, seconds_target=12.0):
    """The reference's own CPU implementation of the path (oracle/_ref, else the pinned port), frame-parallel
    over `threads` host threads like the reference's one-render-thread-per-client model, on the same 4K ->
    320x96 truecolor half-block call.  Returns (Mpix/s nominal, kind, cores, sample description, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_bind as ob
    ring = 8
    frames = np.stack([ob.gen("noise", SRC_W, SRC_H, i) for i in range(ring)])
    u8p = C.POINTER(C.c_uint8)
    R = ob.ref()
    caps = ob.make_caps(LEVEL, MODE)
    fn = C.cast(R.ascii_convert_with_capabilities, C.c_void_p) if R is not None else None
    kind = "reference" if R is not None else "port"
    pal = ob.PALETTES["standard"].encode()
    nbytes = C.c_uint64(0)

    def run(n):
        return ob.port().orc_bench_convert(frames.ctypes.data_as(u8p), ring, SRC_W, SRC_H, COLS, ROWS, LEVEL, MODE,
                                           pal, ob.SCALE_NN, n, threads, fn, C.byref(caps) if fn else None,
                                           C.byref(nbytes))
    run(threads * 2)  # warm tables / caches
    t = run(threads * 4)
    per = t / (threads * 4)
    n = max(threads * 4, int(seconds_target / max(per, 1e-6)))
    t = run(n)
    return n * MPIX / t, kind, threads, "%d renders of a ring of %d LCG-noise 4K frames, %d threads, %.1f s" % (
        n, ring, threads, t), t, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--e2e-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = os.cpu_count() or 1
    config = {"workload": "C3/C5: 3840x2160 RGB24 -> 320x96 truecolor half-block, box-filter downscale, "
                          "ring of %d resident frames per GPU" % args.ring,
              "src": [SRC_W, SRC_H], "cells": [COLS, ROWS], "color": "truecolor", "render_mode": "half-block",
              "palette": "standard", "downscale": "box", "frames_per_step_per_gpu": args.ring,
              "l2": "inputs (%.2f GB per pass) exceed L2; no flush needed" % (args.ring * FRAME_BYTES / 1e9),
              "parallelism": "frames sharded across %d GPU(s), no data-path collective" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        per_step = max(4.0, min(20.0, 120.0 / (steps + args.warmup)))
        vals = []
        for i in range(args.warmup + steps):
            v, kind, cores, sample, secs, n = cpu_reference_leg(ncores, seconds_target=per_step if i >= args.warmup else 2.0)
            if i >= args.warmup:
                vals.append((v, secs, n))
        tot_n = sum(x[2] for x in vals)
        tot_s = sum(x[1] for x in vals)
        value = tot_n * MPIX / tot_s
        cfg = dict(config)
        cfg["downscale"] = "nearest-neighbour (the reference has no box filter; its own path samples 1 px per cell, " \
                           "so Mpix/s is nominal = source pixels / time)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic (LCG noise RGB24)", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": kind,
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import ascii_chat_b200 as acb

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    assert acb.lib().acb200_init(local_rank) == 0, acb.last_error()

    n = args.ring
    cfg = acb.make_cfg(SRC_W, SRC_H, COLS, ROWS * 2, LEVEL, MODE, "standard", scale=acb.SCALE_BOX)
    cap = acb.frame_capacity(cfg)
    g = torch.Generator(device="cuda")
    g.manual_seed(12345 + rank)
    d_in = torch.randint(0, 256, (n, SRC_H, SRC_W, 3), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
    targs = (cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident (HBM -> HBM) throughput: device-timed with CUDA events on the launch stream
    acb.time_batch_device(*targs, max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = acb.launch_count()
    ms_total, ms_kernel = acb.time_batch_device(*targs, args.steps)
    launches = acb.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max, ms_kernel_max = float(t[0]), float(t[1])
    out_bytes = int(d_len.sum().item())

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D + render + D2H + malloc'd strings
    eb = max(1, min(args.e2e_batch, n))
    h_in = torch.empty((eb, SRC_H, SRC_W, 3), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in[:eb])
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * eb)(*[h_in[i].data_ptr() for i in range(eb)])
    outs = (C.c_void_p * eb)()
    lens = (C.c_size_t * eb)()

    def e2e_step():
        rc = acb.render_batch_host_ptrs(cfg, ptrs, eb, outs, lens)
        assert rc == 0, acb.last_error()
        b = sum(lens[i] for i in range(eb))
        acb.free_strings(outs, eb)
        return b

    for _ in range(max(1, args.warmup)):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_val = world * e2e_steps * eb * MPIX / e2e_s

    # ---- the drop-in call in its reference-exact (nearest-neighbour) mode, one frame per call, host buffers
    caps = acb.make_caps(LEVEL, MODE)
    fr0 = h_in[0].numpy()
    for _ in range(3):
        acb.ascii_convert_with_capabilities(fr0, COLS, ROWS, caps, False, False, "standard")
    t0 = time.perf_counter()
    nn_calls = 50
    for i in range(nn_calls):
        acb.ascii_convert_with_capabilities(h_in[i % eb].numpy(), COLS, ROWS, caps, False, False, "standard")
    nn_s = time.perf_counter() - t0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    alg_bytes_launch = n * FRAME_BYTES  # 3 B per source pixel, x frames per launch
    achieved = alg_bytes_launch / (ms_kernel_max / args.steps * 1e-3) / 1e9
    value = world * args.steps * n * MPIX / (ms_total_max * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic (uniform-noise RGB24, worst case "
        "for run-length: ~1.18 MB of ANSI per frame)", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "kernel": "k_render_rows<EM_HB_TRUE, SP_BOX_STREAM>",
                     "algorithmic_bytes_per_launch": alg_bytes_launch,
                     "kernel_ms_per_launch": ms_kernel_max / args.steps,
                     "output_bytes_per_launch": out_bytes, "traffic": traffic_from_profiles()},
        "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": eb * FRAME_BYTES,
                "d2h_bytes_per_step": int(d2h) + 4 * eb, "frames_per_step": eb, "steps": e2e_steps,
                "api": "acb200_render_batch_host (pinned host RGB24 in, malloc'd strings out), box-filter mode"},
        "e2e_dropin_nn": {"value": nn_calls * MPIX / nn_s, "unit": "Mpix/s (nominal)", "ms_per_call": 1e3 * nn_s / nn_calls,
                          "api": "ascii_convert_with_capabilities, one frame per call, reference-exact NN mode, "
                                 "host image in, malloc'd string out"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        v, kind, cores, sample, _, _ = cpu_reference_leg(ncores)
        line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample,
                                "note": "reference path is nearest-neighbour: Mpix/s nominal (source px / time)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
