#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mpixels/s through the fused
downscale -> luminance -> quantise -> glyph kernel at 3840x2160, % of the HBM roofline, 1/2/4/8 GPUs,
with the reference's CPU path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

Workload (config.workload): BASELINE configs[2]/[4] — 3840x2160 RGB24 -> 320x96 truecolor half-block cells.
  value     resident path: box-filter downscale (every source pixel read once: 3 B/px algorithmic), a ring of 256
            distinct frames per GPU already in HBM (6.37 GB per pass >> 126 MB L2, so no L2 flush is needed).
            A step = one pass of the render path over the 256-frame batch, device-timed with CUDA events.
  e2e       the reference-facing call ascii_convert_with_capabilities() (reference-exact nearest-neighbour mode:
            same bytes as the reference) from HOST frames, driven by the same neutral pthread harness as the
            reference arm (bench_harness/caller_threads.c: one caller thread per client, like src/server/render.c).
  e2e_box   same call with the library's downscale switched to the box filter: full frames cross PCIe.
Frames are independent, so ranks take disjoint rings with no data-path collective ("scaling": "weak").
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
HOST_RING = 16
FRAME_BYTES = SRC_W * SRC_H * 3
MPIX = SRC_W * SRC_H / 1e6
METRIC = "Mpixels/s fused RGB->glyph render at 4K"
PALETTE = b"   ...',;:clodxkO0KXNWM"


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
            return json.load(open(p)).get("render_rows_4k_hb_box_bytes_per_256_frame_launch")
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
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)  # let the sampler come up before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()  # the exact PID we started
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


# ------------------------------------------------------------------------------------------- shared harness
def load_harness():
    so = os.path.join(ROOT, "bench_harness", "libcaller_threads.so")
    src = os.path.join(ROOT, "bench_harness", "caller_threads.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "bench_harness")], check=True, stdout=subprocess.DEVNULL)
    H = C.CDLL(so)
    H.harness_run.restype = C.c_double
    H.harness_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_void_p,
                              C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    H.harness_ring_fingerprint.restype = C.c_uint64
    H.harness_ring_fingerprint.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long,
                                           C.c_void_p, C.c_char_p]
    return H


def host_ring(n=HOST_RING, seed=12345):
    """deterministic uniform-noise RGB24 frames in ordinary (pageable) host memory, identical in both arms"""
    import numpy as np
    return np.random.default_rng(seed).integers(0, 256, (n, SRC_H, SRC_W, 3), dtype=np.uint8)


def run_callers(H, fn_ptr, frames, caps, threads, seconds_target, warm_calls=None):
    """calibrate, then time `calls` renders; returns dict(seconds, calls, bytes, failures)"""
    ring = frames.shape[0]
    nb, nf = C.c_uint64(0), C.c_uint64(0)

    def run(calls):
        return H.harness_run(fn_ptr, frames.ctypes.data, ring, SRC_W, SRC_H, COLS, ROWS, C.byref(caps), PALETTE, calls,
                             threads, C.byref(nb), C.byref(nf))
    w = warm_calls or threads * 2
    run(w)
    t = run(w)
    calls = max(threads * 2, int(seconds_target / max(t / w, 1e-7)))
    t = run(calls)
    return {"seconds": t, "calls": calls, "bytes": nb.value, "failures": nf.value}


def reference_entry():
    """(fn pointer, caps struct, kind) of the reference's own CPU implementation: oracle/_ref, else the port shim"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    R = ob.ref()
    if R is None:
        return None, None, "port", ob
    return C.cast(R.ascii_convert_with_capabilities, C.c_void_p), ob.make_caps(LEVEL, MODE), "reference", ob


def cpu_reference_leg(H, frames, threads, seconds_target):
    """The reference's own CPU path (its nearest-neighbour resize + half-block printer), frame-parallel over
    `threads` caller threads.  Falls back to the pinned port only if oracle/_ref was never built."""
    fn, caps, kind, ob = reference_entry()
    if fn is None:  # port: time it through its own helper (kind = "port")
        u8p = C.POINTER(C.c_uint8)
        nb = C.c_uint64(0)
        calls = max(threads * 4, 64)
        t = ob.port().orc_bench_convert(frames.ctypes.data_as(u8p), frames.shape[0], SRC_W, SRC_H, COLS, ROWS, LEVEL,
                                        MODE, PALETTE, ob.SCALE_NN, calls, threads, None, None, C.byref(nb))
        return {"seconds": t, "calls": calls, "bytes": nb.value, "failures": 0}, kind, None
    r = run_callers(H, fn, frames, caps, threads, seconds_target)
    fp = H.harness_ring_fingerprint(fn, frames.ctypes.data, frames.shape[0], SRC_W, SRC_H, COLS, ROWS, C.byref(caps),
                                    PALETTE)
    return r, kind, fp


def server_path_leg(acb, with_reference):
    """the server's per-client entry (stream.c:958-1191) with the senders' frames resident in
    HBM — 9 clients sending 720p, each receiving client a 240x67 truecolor half-block terminal, one render thread
    per receiving client as in src/server/render.c.  The reference leg is its own create_mixed_ascii_frame_for_client
    inside oracle/_ref (single thread: it works on the server's global client table)."""
    import threading
    import numpy as np
    n, sw, sh, W, H = 9, 1280, 720, 240, 67
    rng = np.random.default_rng(777)
    srcs = [rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8) for _ in range(n)]
    for rnd in range(3):  # the first rounds size the slots' double buffers and the staging; time the steady state
        t0 = time.perf_counter()
        for i, s in enumerate(srcs):
            if acb.source_update(i, s) != 0:
                return {"error": str(acb.last_error())}
        upd = (time.perf_counter() - t0) / n
    caps = acb.make_caps(LEVEL, MODE, True)
    slots = list(range(n))
    first = acb.mixed_frame(slots, W, H, caps, "standard")
    per = 150

    def render():
        for _ in range(per):
            acb.mixed_frame(slots, W, H, caps, "standard")
    render()
    t0 = time.perf_counter()
    render()
    one = (time.perf_counter() - t0) / per
    ts = [threading.Thread(target=render) for _ in range(n)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    allc = time.perf_counter() - t0
    out = {"workload": "9 senders x 1280x720 RGB24 resident, 9 receiving clients x 240x67 truecolor half-block, padded",
           "api": "acb200_source_update() per received frame + acb200_mixed_frame() per (client, output frame)",
           "frames_per_s_9_render_threads": n * per / allc, "ms_per_frame_single_thread": one * 1e3,
           "ms_per_source_update": upd * 1e3, "frame_bytes": first[1]}
    if with_reference:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind as ob
        fn = ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame
        exp = fn(srcs, W, H, LEVEL, MODE, "standard", True)
        R = 8
        t0 = time.perf_counter()
        for _ in range(R):
            fn(srcs, W, H, LEVEL, MODE, "standard", True)
        ref = (time.perf_counter() - t0) / R
        out.update({"bytes_identical_to_reference": bool(exp == first), "reference_ms_per_frame_single_thread": ref * 1e3,
                    "reference_kind": "reference" if ob.ref() is not None else "port"})
    for i in range(n):
        acb.source_clear(i)
    return out


def device_extras_leg(acb, torch, d_out, cap, d_len, n, peak):
    """SURVEY.md §8f rows 1 and 4 on resident data, device-timed with CUDA events on torch's current stream (the
    kernels are launched on that stream): the whole-image colour filter (apply_color_filter as an in-place map,
    3 B read + 3 B written per pixel) and the CRC32-C + packet-header scan over the batch's finished strings."""
    ts = torch.cuda.Stream()  # an explicit (non-NULL) stream: the library launches on it, the events are recorded on it
    st = ts.cuda_stream
    assert st, "need a non-default stream handle"
    out = {}

    def timed(fn, iters, warm=3):
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

    d_hdr = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: acb.frame_packets_device(d_out.data_ptr(), cap, d_len.data_ptr(), n, COLS, ROWS,
                                                d_hdr.data_ptr(), st), 10)
    sbytes = int(d_len.sum().item())
    out["frame_packets"] = {"kernels": "k_crc32c_chunks + k_crc32c_finish", "frames": n, "string_bytes": sbytes,
                            "ms_per_batch": ms, "GBs_of_string_bytes": sbytes / (ms * 1e-3) / 1e9,
                            "note": "CRC32-C (lib/network/crc32.c) + 24-byte ascii_frame_packet_t per frame "
                                    "(acip/server.c:203-214)"}
    k = 64
    img = torch.randint(0, 256, (k * SRC_H, SRC_W, 3), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: acb.color_filter_device(img.data_ptr(), SRC_W, k * SRC_H, SRC_W * 3, 3, 0.0, st), 10)
    traffic = 2 * k * FRAME_BYTES
    out["color_filter"] = {"kernel": "k_color_filter", "frames": k, "ms_per_batch": ms, "bound": "hbm",
                           "algorithmic_bytes": traffic, "achieved": traffic / (ms * 1e-3) / 1e9, "unit": "GB/s",
                           "peak": peak, "frac": traffic / (ms * 1e-3) / 1e9 / peak,
                           "Mpix_s": k * MPIX / (ms * 1e-3),
                           "note": "apply_color_filter (color_filter.c:274-346) in place on %d resident 4K frames: "
                                   "3 B read + 3 B written per pixel" % k}
    del img, d_hdr
    return out


def display_path_leg(acb, with_reference):
    """SURVEY.md §8f rows 1+3: the client's display conversion (display.c:484-671) — flip X, green colour filter,
    4K -> 320x96 truecolor half-block — as ONE call from a host frame; the reference leg runs its own functions in
    display.c's order (copy+flip, copy+apply_color_filter, ascii_convert_with_capabilities)."""
    import numpy as np
    img = np.random.default_rng(4242).integers(0, 256, (SRC_H, SRC_W, 3), dtype=np.uint8)
    caps = acb.make_caps(LEVEL, MODE)
    args = (COLS, ROWS, caps, False, False, "standard", True, False, 3, 0.0)
    first = acb.display_convert(img, *args)
    if first is None:
        return {"error": str(acb.last_error())}
    R = 200
    for _ in range(20):
        acb.display_convert(img, *args)
    t0 = time.perf_counter()
    for _ in range(R):
        acb.display_convert(img, *args)
    ours = (time.perf_counter() - t0) / R
    out = {"workload": "one 3840x2160 host frame, flip_x + green filter -> 320x96 truecolor half-block",
           "api": "acb200_display_convert()", "ms_per_call": ours * 1e3, "frame_bytes": len(first)}
    if with_reference:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_bind as ob
        fn = ob.ref_display_convert if ob.ref() is not None else ob.port_display_convert
        kw = dict(cols=COLS, rows=ROWS, level=LEVEL, mode=MODE, palette="standard", flip_x=True, color_filter=3)
        exp = fn(img, **kw)
        t0 = time.perf_counter()
        for _ in range(5):
            fn(img, **kw)
        out.update({"bytes_identical_to_reference": bool(exp == first),
                    "reference_ms_per_call": (time.perf_counter() - t0) / 5 * 1e3,
                    "reference_kind": "reference" if ob.ref() is not None else "port"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="profiling aid: only the device-timed resident loop")
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
        H = load_harness()
        frames = host_ring()
        steps = max(1, args.steps)
        per_step = max(1.0, min(10.0, 100.0 / (steps + args.warmup)))
        tot_s = tot_n = 0
        kind = "reference"
        for i in range(args.warmup + steps):
            r, kind, _ = cpu_reference_leg(H, frames, ncores, per_step if i >= args.warmup else 0.5)
            if i >= args.warmup:
                tot_s += r["seconds"]
                tot_n += r["calls"]
        value = tot_n * MPIX / tot_s
        cfg = dict(config)
        cfg["downscale"] = ("nearest-neighbour: the reference has no box filter; its own path samples 1 px per cell, "
                            "so its Mpix/s is nominal (source pixels / time)")
        sample = "%d calls of ascii_convert_with_capabilities over a ring of %d uniform-noise 4K frames, %d caller " \
                 "threads, %.1f s" % (tot_n, frames.shape[0], ncores, tot_s)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic (uniform-noise RGB24)", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": ncores, "kind": kind,
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    import torch
    import ascii_chat_b200 as acb

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    assert acb.lib().acb200_init(local_rank) == 0, acb.last_error()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor(x, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # ---- resident (HBM -> HBM) throughput: device-timed with CUDA events on the launch stream
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
    acb.time_batch_device(*targs, max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = acb.launch_count()
    ms_total, ms_kernel = acb.time_batch_device(*targs, args.steps)
    launches = acb.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total_max, ms_kernel_max = max_over_ranks([ms_total, ms_kernel])
    out_bytes = int(d_len.sum().item())
    extras = None
    if rank == 0 and world == 1 and not args.resident_only:
        extras = device_extras_leg(acb, torch, d_out, cap, d_len, n, peaks()[0])
    del d_in, d_out, d_scr
    torch.cuda.empty_cache()

    if args.resident_only:
        if rank == 0:
            print(json.dumps({"resident_only": True, "ms_per_step": ms_total_max / args.steps,
                              "kernel_ms_per_launch": ms_kernel_max / args.steps, "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: the reference-facing call from host frames, T caller threads (one per "client")
    H = load_harness()
    frames = host_ring()
    threads = max(4, min(32, ncores // max(1, world)))  # MAX_CLIENTS is 32 (include/ascii-chat/common/limits.h:26)
    caps = acb.make_caps(LEVEL, MODE)
    fn = C.cast(acb.lib().ascii_convert_with_capabilities, C.c_void_p)
    acb.lib().acb200_set_default_scale(acb.SCALE_NN)
    barrier()
    e_nn = run_callers(H, fn, frames, caps, threads, 4.0)
    barrier()
    fp_nn = H.harness_ring_fingerprint(fn, frames.ctypes.data, frames.shape[0], SRC_W, SRC_H, COLS, ROWS,
                                       C.byref(caps), PALETTE)
    acb.lib().acb200_set_default_scale(acb.SCALE_BOX)
    barrier()
    e_box = run_callers(H, fn, frames, caps, threads, 4.0)
    barrier()
    acb.lib().acb200_set_default_scale(acb.SCALE_NN)
    nn_s, box_s = max_over_ranks([e_nn["seconds"], e_box["seconds"]])
    calls_nn, calls_box = e_nn["calls"], e_box["calls"]  # same calibration on every rank is not guaranteed: sum them
    tc = torch.tensor([calls_nn, calls_box, e_nn["failures"] + e_box["failures"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    calls_nn_all, calls_box_all, failures = float(tc[0]), float(tc[1]), int(tc[2])

    # single-caller latency of the drop-in call (what one render thread sees)
    one = run_callers(H, fn, frames, caps, 1, 1.0, warm_calls=8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    alg_bytes_launch = n * FRAME_BYTES  # 3 B per source pixel, x frames per launch
    achieved = alg_bytes_launch / (ms_kernel_max / args.steps * 1e-3) / 1e9
    value = world * args.steps * n * MPIX / (ms_total_max * 1e-3)
    gathered = ROWS * 2 * SRC_W * 3  # NN mode moves only the 192 sampled source rows per frame
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic (uniform-noise RGB24, worst case "
        "for run-length: ~1.18 MB of ANSI per frame)", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "kernel": "k_render_rows_ws2<EM_HB_TRUE> (role-split persistent, direct output)",
                     "algorithmic_bytes_per_launch": alg_bytes_launch,
                     "kernel_ms_per_launch": ms_kernel_max / args.steps,
                     "output_bytes_per_launch": out_bytes,
                     "achieved_incl_output_GBs": (alg_bytes_launch + out_bytes) / (ms_kernel_max / args.steps * 1e-3) / 1e9,
                     "traffic": traffic_from_profiles()},
        "e2e": {"value": calls_nn_all * MPIX / nn_s, "unit": "Mpix/s", "h2d_bytes_per_step": gathered,
                "d2h_bytes_per_step": int(e_nn["bytes"] / max(1, e_nn["calls"])) + 4, "step": "one frame through the call",
                "calls": int(calls_nn_all), "seconds": nn_s, "caller_threads_per_gpu": threads, "failures": failures,
                "api": "ascii_convert_with_capabilities() — reference-exact nearest-neighbour mode, pageable host "
                       "RGB24 in, malloc'd string out, same pthread harness as the reference arm",
                "ring_fingerprint": "%016x" % fp_nn},
        "e2e_box": {"value": calls_box_all * MPIX / box_s, "unit": "Mpix/s", "h2d_bytes_per_step": FRAME_BYTES,
                    "d2h_bytes_per_step": int(e_box["bytes"] / max(1, e_box["calls"])) + 4, "calls": int(calls_box_all),
                    "seconds": box_s, "api": "same call, acb200_set_default_scale(ACB200_SCALE_BOX): whole frames "
                                             "cross PCIe (24.9 MB per frame)"},
        "dropin_single_caller": {"ms_per_call": 1e3 * one["seconds"] / one["calls"],
                                 "mpix_s": one["calls"] * MPIX / one["seconds"]},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        r, kind, fp_ref = cpu_reference_leg(H, frames, ncores, 12.0)
        line["cpu_baseline"] = {"value": r["calls"] * MPIX / r["seconds"], "unit": "Mpix/s", "cores": ncores,
                                "kind": kind,
                                "sample": "%d calls over a ring of %d uniform-noise 4K frames, %d caller threads, %.1f s"
                                          % (r["calls"], frames.shape[0], ncores, r["seconds"]),
                                "note": "reference path is nearest-neighbour: Mpix/s nominal (source px / time)"}
        if fp_ref is not None:
            line["e2e"]["bytes_identical_to_cpu_baseline"] = bool(fp_ref == fp_nn)
    if world == 1:
        line["server_path"] = server_path_leg(acb, not args.no_cpu_baseline)
        line["display_path"] = display_path_leg(acb, not args.no_cpu_baseline)
        if extras:
            line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
