#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: Mpixels/s through the fused
downscale -> luminance -> quantise -> glyph kernel at 3840x2160, % of the HBM roofline, 1/2/4/8 GPUs,
with the reference's CPU path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)

Workload (config.workload): BASELINE configs[2]/[4] — 3840x2160 RGB24 -> 320x96 truecolor half-block cells.
  value      resident path: box-filter downscale (every source pixel read once: 3 B/px algorithmic), a ring of 256
             distinct frames per GPU already in HBM (6.37 GB per pass >> 126 MB L2, so no L2 flush is needed).
             A step = one pass of the render path over the 256-frame batch, device-timed with CUDA events.
  sustained  BASELINE config 5: 100 000 frame-renders (391 passes over the ring), same kernel, clocks sampled inside.
  e2e        the reference-facing call ascii_convert_with_capabilities() (reference-exact nearest-neighbour mode:
             same bytes as the reference) from HOST frames, driven by the same neutral pthread harness as the
             reference arm (bench_harness/caller_threads.c: one caller thread per client, like src/server/render.c),
             ONE process driving all N GPUs through the C ABI (acb200_init_devices) — the server's own structure.
             e2e_per_rank_processes: the same call with one process per GPU (N > 1 only).
  e2e_box    same call with the library's downscale switched to the box filter: full frames cross PCIe (N = 1).
  c4         BASELINE config 4 at every N: 8 clients x 1080p -> 160x48 ANSI-256 -> 320x96 grid, text-space
             (ascii_create_grid) and pixel-space (the server compositor), sharded over the GPUs — NCCL gather between
             rank processes, and peer stores/loads over NVLink inside one process.
Frames are independent, so ranks take disjoint rings with no data-path collective ("scaling": "weak").
One JSON line is printed by rank 0.  See DESIGN.md §6 for what each field means.
"""
import argparse
import ctypes as C
import datetime
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
C5_FRAMES = 100000  # BASELINE config 5
CALLERS_PER_GPU = 16  # knee of the caller sweep on one PCIe link (profiles/r02a_e2e_sweep_pixels.txt)


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
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


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


def pick_threads(H, fn_ptr, frames, caps, candidates, seconds=0.6):
    """caller-thread count at the knee: probe each candidate briefly (untimed region), keep the fastest.  The knee moves
    with the box (cores per GPU, PCIe links in use): 16 callers on 16 cores and one GPU, 16 on 24 cores and two GPUs
    (profiles/r02a_e2e_sweep_pixels.txt, r02b_e2e_inproc2.txt), and a caller per core is past it."""
    best, best_fps, probes = None, 0.0, {}
    for t in sorted(set(int(c) for c in candidates if c >= 1)):
        r = run_callers(H, fn_ptr, frames, caps, t, seconds)
        fps = r["calls"] / r["seconds"]
        probes[t] = round(fps)
        if fps > best_fps:
            best, best_fps = t, fps
    return best, probes


def thread_candidates(ncores, n_gpus):
    cap = CALLERS_PER_GPU * n_gpus
    return [min(cap, c) for c in (max(2, ncores // 2), max(2, (2 * ncores) // 3), max(2, ncores - 2), ncores,
                                  (3 * ncores) // 2)]


SYNC_NAMES = {0: "spin", 1: "block", 2: "hybrid", 3: "yield"}


def pick_config(lib, H, fn_ptr, frames, caps, combos, seconds=0.5):
    """probe (caller threads, fetch depth, wait mode) combinations briefly (untimed region), keep the fastest"""
    best, probes = None, []
    for (t, depth, sync) in combos:
        lib.acb200_set_fetch_depth(depth)
        lib.acb200_set_sync_mode(sync, 30)
        r = run_callers(H, fn_ptr, frames, caps, t, seconds)
        fps = r["calls"] / r["seconds"]
        probes.append({"threads": t, "fetch_depth": depth, "wait": SYNC_NAMES[sync], "fps": round(fps)})
        if best is None or fps > best[0]:
            best = (fps, t, depth, sync)
    return best[1:], probes


def e2e_api_legs(acb, H, frames, caps, n_gpus, seconds):
    """The drop-in call end to end, two ways.  `pageable`: frames in ordinary memory — the calling thread gathers the
    sampled pixels (0.18 MB per 4K frame cross PCIe).  `registered`: the SAME frames page-locked once, untimed
    (acb200_register_host_memory — what a server does with its per-client frame buffers): a call may then let the
    device fetch the sampled rows (2.2 MB per frame) and spend no core time on the input at all.  Caller threads, wait
    mode and — registered — whether the fetch is used are probed per box: hosts with few cores per GPU win with the
    fetch, hosts with many do not (profiles/r02n_e2e_registered*.txt)."""
    lib = acb.lib()
    fn = C.cast(lib.ascii_convert_with_capabilities, C.c_void_p)
    lib.acb200_set_default_scale(acb.SCALE_NN)
    ncores = os.cpu_count() or 1
    cand = sorted(set(thread_candidates(ncores, n_gpus)))

    def measure(cfg3, probes):
        t, depth, sync = cfg3
        lib.acb200_set_fetch_depth(depth)
        lib.acb200_set_sync_mode(sync, 30)
        ph = (C.c_uint64 * 5)()
        lib.acb200_host_phase_stats(ph, 1)
        r = run_callers(H, fn, frames, caps, t, seconds)
        lib.acb200_host_phase_stats(ph, 1)
        fp = H.harness_ring_fingerprint(fn, frames.ctypes.data, frames.shape[0], SRC_W, SRC_H, COLS, ROWS,
                                        C.byref(caps), PALETTE)
        r.update({"threads": t, "fetch_depth": depth, "wait": SYNC_NAMES[sync], "probes": probes, "n_gpus": n_gpus,
                  "ring_fingerprint": "%016x" % fp,
                  "host_us_per_call": {k: round(ph[i] / max(1, ph[4]) / 1e3, 1)
                                       for i, k in enumerate(("gather", "enqueue", "wait", "copy_out"))}})
        return r

    # pageable frames: caller count first (spinning wait), then the yielding wait at the best count
    (t0, _, _), p1 = pick_config(lib, H, fn, frames, caps, [(t, 0, 0) for t in cand])
    best_pg, p2 = pick_config(lib, H, fn, frames, caps, [(t0, 0, 0), (t0, 0, 3), (min(2 * t0, 2 * ncores), 0, 3)])
    pageable = measure(best_pg, p1 + p2)
    one = run_callers(H, fn, frames, caps, 1, 1.0, warm_calls=8)
    pageable["single_caller_ms"] = 1e3 * one["seconds"] / one["calls"]
    # the same frames page-locked
    registered = None
    if lib.acb200_register_host_memory(frames.ctypes.data, frames.nbytes) == 0:
        # fetching callers mostly wait (the link is the bound: ~8-16 of them per GPU saturate it), so more callers than
        # cores make sense — with the yielding wait; spinning ones only up to one per core
        fetch_t = sorted(set(max(2, c) for c in (ncores // 2, ncores, (3 * ncores) // 2, 2 * ncores, 3 * ncores,
                                                 4 * ncores) if c <= 16 * n_gpus or c <= ncores))
        best_rg, p3 = pick_config(lib, H, fn, frames, caps,
                                  [best_pg] + [(t, -1, 0) for t in fetch_t if t <= ncores] +
                                  [(t, -1, 3) for t in fetch_t])
        if best_rg == best_pg:
            # the probe found nothing faster than the host-gathered configuration: page-locked or not, it is the same
            # code path on the same frames, so the pageable measurement stands for both (no second, noisier sample)
            registered = dict(pageable)
            registered["probes"] = p3
            registered["same_as_pageable"] = True
        else:
            registered = measure(best_rg, p3)
        lib.acb200_set_fetch_depth(-1)
        lib.acb200_set_sync_mode(0, 30)
        one = run_callers(H, fn, frames, caps, 1, 1.0, warm_calls=8)
        registered["single_caller_ms_fetch"] = 1e3 * one["seconds"] / one["calls"]
        lib.acb200_unregister_host_memory(frames.ctypes.data)
    else:
        acb.last_error()
    lib.acb200_set_fetch_depth(0)
    lib.acb200_set_sync_mode(0, 30)
    return {"pageable": pageable, "registered": registered, "launches": acb.launch_count()}


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


def cpu_box_leg(frames, threads, seconds_target):
    """Same-work CPU baseline for `value`: the box filter of DESIGN.md §3 on the CPU
    (oracle's column-sum arrangement, AVX2) + the compiled reference's own printer on the filtered image."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    u8p = C.POINTER(C.c_uint8)
    R = ob.ref()
    fn = C.cast(R.image_print_with_capabilities, C.c_void_p) if R is not None else None
    caps = ob.make_caps(LEVEL, MODE)
    nb = C.c_uint64(0)

    def run(calls):
        return ob.port().orc_bench_box(frames.ctypes.data_as(u8p), frames.shape[0], SRC_W, SRC_H, COLS, ROWS * 2, LEVEL,
                                       MODE, PALETTE, calls, threads, fn, C.byref(caps) if fn else None, C.byref(nb))
    w = threads * 2
    t = run(w)
    calls = max(w, int(seconds_target / max(t / w, 1e-6)))
    t = run(calls)
    return {"value": calls * MPIX / t, "unit": "Mpix/s", "cores": threads, "calls": calls, "seconds": t,
            "kind": "port box filter + %s printer" % ("reference" if fn else "port")}


def host_bytes_per_frame(out_bytes):
    """host DRAM traffic one e2e frame causes: 64-byte lines of the source the gather touches, the
    staged pixels written and DMA-read, the string DMA-written into pinned memory, then copied into malloc'd memory"""
    xr = ((SRC_W << 16) // COLS) + 1
    yr = ((SRC_H << 16) // (ROWS * 2)) + 1
    lines = set()
    R = SRC_W * 3
    for y in range(ROWS * 2):
        sy = min((y * yr) >> 16, SRC_H - 1)
        for x in range(COLS):
            o = sy * R + min((x * xr) >> 16, SRC_W - 1) * 3
            lines.add(o >> 6)
            lines.add((o + 2) >> 6)
    staged = COLS * ROWS * 2 * 3
    return {"source_lines_read": len(lines) * 64, "staging_written": staged, "h2d_dma_read": staged,
            "d2h_dma_written": out_bytes, "string_copy_read": out_bytes, "string_copy_written": out_bytes,
            "total": len(lines) * 64 + 2 * staged + 3 * out_bytes}


# ------------------------------------------------------------------------------------------- legs
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
    upd = upd_pinned = 0.0
    for rnd in range(3):  # the first rounds size the slots' double buffers and the staging; time the steady state
        t0 = time.perf_counter()
        for i, s in enumerate(srcs):
            if acb.source_update(i, s) != 0:
                return {"error": str(acb.last_error())}
        upd = (time.perf_counter() - t0) / n
    for i in range(n):  # the receive-into-pinned form: the transport's buffer is the slot's registered one
        acb.lib().acb200_source_acquire(i, srcs[i].nbytes)
    for rnd in range(3):
        t0 = time.perf_counter()
        for i, s in enumerate(srcs):
            acb.lib().acb200_source_commit(i, sw, sh)
        upd_pinned = (time.perf_counter() - t0) / n
    for i, s in enumerate(srcs):
        acb.source_update(i, s)
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
           "ms_per_source_update": upd * 1e3, "ms_per_source_commit_pinned": upd_pinned * 1e3,
           "frame_bytes": first[1]}
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


def crc_overlap_leg(acb, torch, cfg, d_in, n, cap):
    """the CRC32-C scan is bound by its byte recurrence, not by HBM, so it hides behind the NEXT
    batch's render when the two run on different streams (double-buffered arenas): K batches rendered + packaged,
    serial on one stream vs overlapped on two.  Event-timed on the render stream."""
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    bufs = []
    for _ in range(2):
        bufs.append((torch.empty(n * cap, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
                     torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda"),
                     torch.zeros(n * 24, dtype=torch.uint8, device="cuda")))
    K = 12

    def run(overlap):
        torch.cuda.synchronize()
        crc_done = [None, None]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sa)
        for i in range(K):
            out, ln, scr, hdr = bufs[i & 1]
            if overlap and crc_done[i & 1] is not None:
                sa.wait_event(crc_done[i & 1])  # the arena is free once its previous scan has finished
            acb.render_batch_device(cfg, d_in.data_ptr(), n, out.data_ptr(), cap, ln.data_ptr(), scr.data_ptr(), sa.cuda_stream)
            if overlap:
                ev = torch.cuda.Event()
                ev.record(sa)
                sb.wait_event(ev)
                acb.frame_packets_device(out.data_ptr(), cap, ln.data_ptr(), n, COLS, ROWS, hdr.data_ptr(), sb.cuda_stream)
                crc_done[i & 1] = torch.cuda.Event()
                crc_done[i & 1].record(sb)
            else:
                acb.frame_packets_device(out.data_ptr(), cap, ln.data_ptr(), n, COLS, ROWS, hdr.data_ptr(), sa.cuda_stream)
        if overlap:
            sa.wait_stream(sb)
        e1.record(sa)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    run(False), run(True)
    serial, over = run(False), run(True)
    hdr_ok = bool(torch.equal(bufs[0][3], bufs[1][3]))  # both arenas hold the same batch: same headers
    del bufs
    torch.cuda.empty_cache()
    return {"ms_per_batch_serial": serial, "ms_per_batch_overlapped": over, "batches": K, "frames_per_batch": n,
            "headers_equal": hdr_ok,
            "note": "render (stream A) || CRC32-C + packet headers of the previous batch (stream B), double-buffered arenas"}


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
    out["frame_packets"] = {"kernels": "k_crc32c_plan + k_crc32c_rows + k_crc32c_tail", "frames": n, "string_bytes": sbytes,
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


def resident_variants_leg(acb, torch, peak):
    """the other resident-batch shapes next to the headline (device-timed, 64-frame 4K batches): the reference-exact
    nearest-neighbour mode on noise and on flat frames (latency/output-bound, not HBM-bound), and the box render with
    a colour filter fused in."""
    n = 64
    out = {}
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    noise = torch.randint(0, 256, (n, SRC_H, SRC_W, 3), dtype=torch.uint8, device="cuda", generator=g)
    band = torch.randint(0, 256, (n, SRC_H // 40 + 1, 1, 3), dtype=torch.uint8, device="cuda", generator=g)
    flat = band.repeat_interleave(40, dim=1)[:, :SRC_H].expand(n, SRC_H, SRC_W, 3).contiguous()

    def run(d_in, **kw):
        cfg = acb.make_cfg(SRC_W, SRC_H, COLS, ROWS * 2, LEVEL, MODE, "standard", **kw)
        cap = acb.frame_capacity(cfg)
        d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
        d_len = torch.empty(n, dtype=torch.int32, device="cuda")
        d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
        a = (cfg, d_in.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr())
        acb.time_batch_device(*a, 8)
        tot, ker = acb.time_batch_device(*a, 20)
        return tot / 20, ker / 20
    for name, d_in in (("noise", noise), ("flat", flat)):
        ms, _ = run(d_in, scale=acb.SCALE_NN)
        out["nn_" + name] = {"ms_per_256_frames": ms * 256 / n, "frames_per_s": n / ms * 1e3,
                             "Mcells_per_s": n * COLS * ROWS / ms / 1e3}
    ms, ker = run(noise, scale=acb.SCALE_BOX, color_filter=3)
    gbs = n * FRAME_BYTES / (ker * 1e-3) / 1e9
    out["box_filtered"] = {"filter": "green (apply_color_filter fused into the band sums)", "ms_per_64_frames": ms,
                           "achieved": gbs, "unit": "GB/s", "peak": peak, "frac": gbs / peak}
    del noise, flat, band
    torch.cuda.empty_cache()
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
    # the digital-rain stage on that string (display.c:657-671): state on the GPU, string work on the device
    rain = acb.DigitalRain(COLS, ROWS, 3)
    r_first = rain.apply(first, 0.016)
    t0 = time.perf_counter()
    for _ in range(50):
        rain.apply(first, 0.016)
    out["digital_rain"] = {"api": "digital_rain_apply()", "ms_per_call": (time.perf_counter() - t0) / 50 * 1e3,
                           "in_bytes": len(first), "out_bytes": len(r_first or b"")}
    rain.close()
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
        rr = (ob.RefRain if ob.ref() is not None else ob.PortRain)(COLS, ROWS, 3)
        r_exp = rr.apply(first, 0.016)
        t0 = time.perf_counter()
        for _ in range(5):
            rr.apply(first, 0.016)
        out["digital_rain"].update({"bytes_identical_to_reference": bool(r_exp == r_first),
                                    "reference_ms_per_call": (time.perf_counter() - t0) / 5 * 1e3})
        rr.close()
    return out


# ---- BASELINE config 4: 8 clients x 1080p -> 160x48 ANSI-256 -> 320x96 grid
C4_N, C4_W, C4_H, C4_COLS, C4_ROWS, C4_LEVEL, C4_MODE, C4_GW, C4_GH = 8, 1920, 1080, 160, 48, 2, 0, 320, 96


def c4_sources(ob):
    return [ob.gen(("noise", "bars", "gradient")[c % 3], C4_W, C4_H, c) for c in range(C4_N)]


def c4_expected(ob, srcs, nul):
    """the checker's answers: text grid of the checker's cells, and the server's mixed frame (both viewer sizes)"""
    conv = ob.ref_convert if ob.ref() is not None else ob.port_convert
    grid = ob.ref_create_grid if ob.ref() is not None else ob.port_create_grid
    mixed = ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame
    cells = [conv(s, C4_COLS, C4_ROWS, C4_LEVEL, C4_MODE) for s in srcs]
    g = grid([c + (b"\0" if nul else b"") for c in cells], C4_GW, C4_GH)
    m = {(w, h): mixed(srcs, w, h, C4_LEVEL, C4_MODE, "standard", True)[0] for (w, h) in ((C4_COLS, C4_ROWS), (C4_GW, C4_GH))}
    return g, m, "reference" if ob.ref() is not None else "port"


def c4_nccl_leg(acb, torch, dist, rank, world):
    """one process per GPU: clients sharded c % world, text-space = render + all-gather of the fixed-pitch arenas +
    ascii_create_grid on rank 0 (multi.GridPipeline); pixel-space = NN-resize to the cell images + all-gather +
    composite/convert on rank 0 (multi.PixelGridPipeline).  Wall clock around K steps, max over ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    from ascii_chat_b200 import multi
    srcs = c4_sources(ob)
    mine = multi.shard_indices(C4_N, rank, world)
    d_mine = [torch.from_numpy(srcs[c]).cuda() for c in mine]
    res = {}

    def timed(step, K=200, warm=20):
        out = None
        for _ in range(warm):
            out = step()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            out = step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return out, float(dt) / K
    cfg = acb.make_cfg(C4_W, C4_H, C4_COLS, C4_ROWS, C4_LEVEL, C4_MODE)
    pipe = multi.GridPipeline(acb, cfg, C4_N, C4_GW, C4_GH)
    batch = torch.stack(d_mine).contiguous() if d_mine else None
    g, t_text = timed(lambda: pipe.step(batch))
    pix = {}
    for (w, h) in ((C4_COLS, C4_ROWS), (C4_GW, C4_GH)):
        pp = multi.PixelGridPipeline(acb, [(C4_W, C4_H)] * C4_N, w, h, acb.make_caps(C4_LEVEL, C4_MODE, True), "standard")
        pix[(w, h)] = timed(lambda: pp.step(d_mine))
    if rank == 0:
        eg, em, kind = c4_expected(ob, srcs, nul=False)
        res = {"text_space": {"ms_per_grid": t_text * 1e3, "grids_per_s": 1 / t_text,
                              "collective": "all_gather(fixed-pitch string arenas) + all_gather(lengths)",
                              "bytes_identical": bool(g == eg[0] or g == eg[0][:eg[1]]), "grid_bytes": len(g or b"")},
               "pixel_space": {"%dx%d" % k: {"ms_per_frame": v[1] * 1e3, "frames_per_s": 1 / v[1],
                                             "bytes_identical": bool(v[0] == em[k]), "frame_bytes": len(v[0] or b"")}
                               for k, v in pix.items()},
               "pixel_space_collective": "all_gather(NN-resized cell images, %d B per client)" % (53 * 30 * 3),
               "checker": kind, "transport": "NCCL over NVLink, one process per GPU" if world > 1 else "NCCL, one rank"}
    del d_mine, batch
    return res


def c4_inprocess(n_gpus):
    """ONE process, n_gpus GPUs behind the C ABI: slots sharded c % n_gpus; acb200_grid_frame renders every cell on the
    GPU that owns the client and stores its rows into the composing GPU's arena over NVLink; acb200_mixed_frame reads the
    remote sources in place (peer loads).  Runs in its own process (the pool must be set up before any other call)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as ob
    import ascii_chat_b200 as acb
    assert acb.init_devices(list(range(n_gpus))) == 0, acb.last_error()
    acb.lib().acb200_bind_thread(0)
    srcs = c4_sources(ob)
    for rnd in range(2):
        for i, s in enumerate(srcs):
            assert acb.source_update(i, s) == 0, acb.last_error()
    slots = list(range(C4_N))
    eg, em, kind = c4_expected(ob, srcs, nul=True)

    def timed(fn, K=300, warm=30):
        out = None
        for _ in range(warm):
            out = fn()
        t0 = time.perf_counter()
        for _ in range(K):
            out = fn()
        return out, (time.perf_counter() - t0) / K
    caps_cell = acb.make_caps(C4_LEVEL, C4_MODE)
    g, t_text = timed(lambda: acb.grid_frame(slots, C4_COLS, C4_ROWS, caps_cell, "standard", C4_GW, C4_GH))
    res = {"n_gpus": n_gpus, "slot_devices": [acb.lib().acb200_source_device(i) for i in slots],
           "text_space": {"ms_per_grid": t_text * 1e3, "grids_per_s": 1 / t_text, "api": "acb200_grid_frame()",
                          "bytes_identical": bool(g == (eg[0], eg[1])), "grid_bytes": len(g[0] or b"")},
           "pixel_space": {}, "checker": kind,
           "transport": "peer stores / loads over NVLink inside one process" if n_gpus > 1 else "one GPU"}
    caps_v = acb.make_caps(C4_LEVEL, C4_MODE, True)
    for (w, h) in ((C4_COLS, C4_ROWS), (C4_GW, C4_GH)):
        m, t = timed(lambda: acb.mixed_frame(slots, w, h, caps_v, "standard"))
        res["pixel_space"]["%dx%d" % (w, h)] = {"ms_per_frame": t * 1e3, "frames_per_s": 1 / t,
                                                "api": "acb200_mixed_frame()", "bytes_identical": bool(m[0] == em[(w, h)]),
                                                "frame_bytes": len(m[0] or b"")}
    for i in slots:
        acb.source_clear(i)
    acb.lib().acb200_shutdown()
    return res


def e2e_inprocess(n_gpus, threads, seconds):
    """ONE process, n_gpus GPUs behind the C ABI, caller threads leased round-robin to the GPUs"""
    import ascii_chat_b200 as acb
    assert acb.init_devices(list(range(n_gpus))) == 0, acb.last_error()
    H = load_harness()
    frames = host_ring()
    caps = acb.make_caps(LEVEL, MODE)
    r = e2e_api_legs(acb, H, frames, caps, n_gpus, seconds)
    acb.lib().acb200_shutdown()
    return r


def run_leg_subprocess(leg, n_gpus, extra=()):
    """a leg that needs its own process (its own device pool): python bench.py --leg ... ; returns its JSON"""
    cmd = [sys.executable, os.path.abspath(__file__), "--leg", leg, "--gpus", str(n_gpus)] + list(extra)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE",
                                                             "GROUP_RANK", "ROLE_RANK", "TORCHELASTIC_RUN_ID")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": (r.stderr or r.stdout)[-600:]}
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="profiling aid: only the device-timed resident loop")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--leg", default=None, choices=[None, "e2e-inprocess", "c4-inprocess"], help="internal")
    ap.add_argument("--threads", type=int, default=0, help="internal (with --leg)")
    args = ap.parse_args()

    if args.leg == "e2e-inprocess":
        print(json.dumps(e2e_inprocess(args.gpus, args.threads, 4.0)))
        return
    if args.leg == "c4-inprocess":
        print(json.dumps(c4_inprocess(args.gpus)))
        return

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

    store = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        store = dist.distributed_c10d._get_default_store()
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

    def rank0_alone(key, fn):
        """rank 0 runs fn() while the other ranks SLEEP on the rendezvous store (an NCCL barrier would spin their cores,
        which are the cores the host-side legs are measuring)"""
        if world == 1:
            return fn()
        if rank == 0:
            try:
                return fn()
            finally:
                store.set(key, b"1")
        store.wait([key], datetime.timedelta(seconds=1800))
        return None

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

    # ---- BASELINE config 5: 100 000 frame-renders over the resident ring, one timed region, clocks sampled inside
    sustained = None
    if not args.no_sustained and not args.resident_only:
        passes = (C5_FRAMES + n - 1) // n
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        barrier()
        s_total, s_kernel = acb.time_batch_device(*targs, passes)
        barrier()
        clocks2 = sampler2.stop() if rank == 0 else None
        s_total_max, s_kernel_max = max_over_ranks([s_total, s_kernel])
        sustained = {"frames_per_gpu": passes * n, "passes": passes, "seconds": s_total_max / 1e3,
                     "value": world * passes * n * MPIX / (s_total_max * 1e-3), "unit": "Mpix/s",
                     "kernel_GBs": passes * n * FRAME_BYTES / (s_kernel_max * 1e-3) / 1e9, "clocks": clocks2,
                     "workload": "BASELINE config 5: %d frame-renders per GPU cycling the %d-frame resident ring" % (passes * n, n)}
    extras = None
    if rank == 0 and world == 1 and not args.resident_only:
        extras = device_extras_leg(acb, torch, d_out, cap, d_len, n, peaks()[0])
        del d_out, d_scr
        torch.cuda.empty_cache()
        extras["frame_packets_overlapped"] = crc_overlap_leg(acb, torch, cfg, d_in, n, cap)
        d_out = d_scr = None
    del d_in, d_out, d_scr
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.resident_only:
        extras.update(resident_variants_leg(acb, torch, peaks()[0]))

    if args.resident_only:
        if rank == 0:
            print(json.dumps({"resident_only": True, "ms_per_step": ms_total_max / args.steps,
                              "kernel_ms_per_launch": ms_kernel_max / args.steps, "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the drop-in call
    H = load_harness()
    frames = host_ring()
    caps = acb.make_caps(LEVEL, MODE)
    fn = C.cast(acb.lib().ascii_convert_with_capabilities, C.c_void_p)
    acb.lib().acb200_set_default_scale(acb.SCALE_NN)
    legs = per_rank = e_box = None
    if world == 1:
        legs = e2e_api_legs(acb, H, frames, caps, 1, 4.0)
        acb.lib().acb200_set_default_scale(acb.SCALE_BOX)
        e_box = run_callers(H, fn, frames, caps, legs["pageable"]["threads"], 3.0)
        acb.lib().acb200_set_default_scale(acb.SCALE_NN)
    else:
        # one process per GPU (pageable frames): every rank drives its own GPU with its share of the caller threads
        threads = max(2, min(CALLERS_PER_GPU, ncores // world))
        barrier()
        e_nn = run_callers(H, fn, frames, caps, threads, 4.0)
        barrier()
        fp_nn = H.harness_ring_fingerprint(fn, frames.ctypes.data, frames.shape[0], SRC_W, SRC_H, COLS, ROWS,
                                           C.byref(caps), PALETTE)
        (nn_s,) = max_over_ranks([e_nn["seconds"]])
        tc = torch.tensor([e_nn["calls"], e_nn["failures"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
        per_rank = {"value": float(tc[0]) * MPIX / nn_s, "unit": "Mpix/s", "calls": int(tc[0]), "seconds": nn_s,
                    "caller_threads_per_gpu": threads, "processes": world, "failures": int(tc[1]),
                    "ring_fingerprint": "%016x" % fp_nn, "frames": "pageable"}

    # ---- BASELINE config 4 over NCCL (all ranks), then the single-process legs (rank 0; the others sleep)
    c4 = c4_nccl_leg(acb, torch, dist if world > 1 else _SingleRankDist(torch), rank, world)
    if world > 1: # ONE process drives all the GPUs (a sub-process of rank 0; the other ranks sleep)
        legs = rank0_alone("e2e_inproc", lambda: run_leg_subprocess("e2e-inprocess", world))
    c4_in = rank0_alone("c4_inproc", lambda: run_leg_subprocess("c4-inprocess", world))

    def rank0_tail():
        peak, peak_src = peaks()
        alg_bytes_launch = n * FRAME_BYTES  # 3 B per source pixel, x frames per launch
        achieved = alg_bytes_launch / (ms_kernel_max / args.steps * 1e-3) / 1e9
        value = world * args.steps * n * MPIX / (ms_total_max * 1e-3)
        def e2e_record(r, frames_kind):
            out_per_frame = int(r["bytes"] / max(1, r["calls"]))
            fetched = r["fetch_depth"] != 0
            # the host gather ships the sampled pixels; the device-side fetch reads the sampled ROWS over the link
            h2d = ROWS * 2 * SRC_W * 3 if fetched else COLS * ROWS * 2 * 3
            hb = host_bytes_per_frame(out_per_frame)
            if fetched:  # no core touches the source; the DMA reads the rows
                hb.update({"source_lines_read": 0, "staging_written": 0, "h2d_dma_read": h2d})
                hb["total"] = h2d + 3 * out_per_frame
            return {"value": r["calls"] * MPIX / r["seconds"], "unit": "Mpix/s", "calls": r["calls"],
                    "seconds": r["seconds"], "frames_per_s": r["calls"] / r["seconds"], "caller_threads": r["threads"],
                    "wait": r["wait"], "input": "the device fetches the sampled rows from the page-locked frame" if fetched else "the caller gathers the sampled pixels",
                    "probes": r["probes"], "processes": 1, "failures": r["failures"],
                    "ring_fingerprint": r["ring_fingerprint"], "host_us_per_call": r["host_us_per_call"],
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_per_frame + 4, "host_bytes_per_step": hb,
                    "step": "one frame through the call", "frames": frames_kind,
                    "structure": "one process, one GPU" if world == 1 else
                    "ONE process, %d GPUs behind the C ABI (acb200_init_devices), caller threads leased round-robin" % world,
                    "api": "ascii_convert_with_capabilities() — reference-exact nearest-neighbour mode, host RGB24 in, "
                           "malloc'd string out, same pthread harness as the reference arm"}

        if legs and "error" not in legs:
            e2e_pg = e2e_record(legs["pageable"], "pageable host memory")
            if legs.get("registered"):
                e2e = e2e_record(legs["registered"], "the same frames page-locked once, outside the timed region "
                                                     "(acb200_register_host_memory)")
                e2e["single_caller_ms_fetch"] = legs["registered"].get("single_caller_ms_fetch")
                if legs["registered"].get("same_as_pageable"):
                    e2e["note"] = ("no probed configuration on page-locked frames (device-side fetch, more callers, "
                                   "yielding wait) beat the host-gathered one on this box: the number is the "
                                   "e2e_pageable measurement (same code path, same frames)")
            else:
                e2e = dict(e2e_pg)
                e2e["note"] = "cudaHostRegister of the frame ring failed on this box: pageable frames"
        else:  # the single-process leg failed: fall back to the per-rank-process numbers
            e2e = dict(per_rank or {})
            e2e.update({"h2d_bytes_per_step": COLS * ROWS * 2 * 3, "d2h_bytes_per_step": None,
                        "structure": "one process per GPU (in-process leg failed: %s)" % (legs or {}).get("error")})
            e2e_pg = None
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
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if sustained:
            line["sustained"] = sustained
        if e2e_pg:
            line["e2e_pageable"] = e2e_pg
        if per_rank:
            line["e2e_per_rank_processes"] = per_rank
        if e_box:
            line["e2e_box"] = {"value": e_box["calls"] * MPIX / e_box["seconds"], "unit": "Mpix/s",
                               "h2d_bytes_per_step": FRAME_BYTES,
                               "d2h_bytes_per_step": int(e_box["bytes"] / max(1, e_box["calls"])) + 4,
                               "calls": e_box["calls"], "seconds": e_box["seconds"],
                               "api": "same call, acb200_set_default_scale(ACB200_SCALE_BOX): whole frames cross PCIe "
                                      "(24.9 MB per frame)"}
        if legs and "error" not in legs:
            ms1 = legs["pageable"]["single_caller_ms"]
            line["dropin_single_caller"] = {"ms_per_call": ms1, "mpix_s": MPIX / (ms1 * 1e-3)}
        line["c4"] = {"workload": "BASELINE config 4: %d clients x %dx%d -> %dx%d ANSI-256 -> %dx%d grid; pixel-space: "
                                  "the server compositor for %dx%d and %dx%d viewers" % (
                                      C4_N, C4_W, C4_H, C4_COLS, C4_ROWS, C4_GW, C4_GH, C4_COLS, C4_ROWS, C4_GW, C4_GH),
                      "rank_processes_nccl": c4, "one_process": c4_in}
        if not args.no_cpu_baseline:
            r, kind, fp_ref = cpu_reference_leg(H, frames, ncores, 10.0)
            line["cpu_baseline"] = {"value": r["calls"] * MPIX / r["seconds"], "unit": "Mpix/s", "cores": ncores,
                                    "kind": kind,
                                    "sample": "%d calls over a ring of %d uniform-noise 4K frames, %d caller threads, %.1f s"
                                              % (r["calls"], frames.shape[0], ncores, r["seconds"]),
                                    "note": "reference path is nearest-neighbour: Mpix/s nominal (source px / time)"}
            if fp_ref is not None:
                for k in ("e2e", "e2e_pageable"):
                    if line.get(k) and line[k].get("ring_fingerprint"):
                        line[k]["bytes_identical_to_cpu_baseline"] = bool("%016x" % fp_ref == line[k]["ring_fingerprint"])
            r1, _, _ = cpu_reference_leg(H, frames, 1, 2.0)
            line["cpu_baseline_1thread"] = {"value": r1["calls"] * MPIX / r1["seconds"], "unit": "Mpix/s", "cores": 1,
                                            "kind": kind, "ms_per_frame": 1e3 * r1["seconds"] / r1["calls"]}
            # the same work as `value` (every source pixel read): CPU box filter + the reference's printer
            line["cpu_baseline_box"] = cpu_box_leg(frames, ncores, 5.0)
            line["cpu_baseline_box_1thread"] = cpu_box_leg(frames, 1, 2.0)
            line["same_work_ratio"] = {"value_over_cpu_baseline_box": value / line["cpu_baseline_box"]["value"],
                                       "note": "both arms read all 8.3 M pixels of every frame (box filter)"}
        if world == 1:
            line["server_path"] = server_path_leg(acb, not args.no_cpu_baseline)
            line["display_path"] = display_path_leg(acb, not args.no_cpu_baseline)
            if extras:
                line.update(extras)
        print(json.dumps(line))

    rank0_alone("tail", rank0_tail)
    import torch.distributed as _d
    if _d.is_initialized():
        _d.destroy_process_group()


class _SingleRankDist:
    """world == 1: the NCCL leg's collectives degenerate; torch.distributed is still used (one-rank gloo-free group)"""

    def __init__(self, torch):
        import torch.distributed as dist
        self._d = dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", str(29500 + os.getpid() % 2000))
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
        self.ReduceOp = dist.ReduceOp

    def barrier(self):
        self._d.barrier()

    def all_reduce(self, t, op=None):
        self._d.all_reduce(t, op=op)


if __name__ == "__main__":
    main()
