"""Multi-GPU plumbing for the grid path (BASELINE config 4): one process per GPU, torch.distributed.

The render path shards embarrassingly: client streams (or frames) are independent, so rank r renders the
clients c with c % world == r on its own GPU with no data-path collective.  The ONE real exchange step of
the reference's design is the grid: rendered client frames are gathered to the composing rank, which
builds the final canvas (text-space: ascii_create_grid, lib/video/ascii/ascii.c:602-885; host.c:696-717).

Everything here is backend-agnostic tensor plumbing (NCCL over NVLink on the GPU box, gloo on CPU for the
world_size-2 tests); the rendering and the composition are C-ABI calls into libasciichat_b200.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """client c -> rank c % world"""
    return list(range(rank, n_items, world))


def owner_of(item, world):
    return item % world


def gather_variable(local, n_items, dst=0, device=None, group=None):
    """Gather variable-length byte strings to `dst`.

    local: dict {item_index: 1-D uint8 tensor} for the items this rank owns (sharded with shard_indices).
    Returns on dst a list of n_items uint8 tensors (on `device`), elsewhere None.

    Two collectives, both fixed-shape so they map onto NCCL directly:
      1. all_gather of the per-item lengths (int64, n_items slots, zero where not owned, then summed),
      2. all_gather of each rank's payload packed at fixed pitch = the global max length
         (payloads are <= ~1.2 MB per 4K frame: latency-bound, so one batched call beats per-item sends).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device or (next(iter(local.values())).device if local else torch.device("cpu"))
    lens = torch.zeros(n_items, dtype=torch.int64, device=device)
    for i, t in local.items():
        lens[i] = t.numel()
    dist.all_reduce(lens, op=dist.ReduceOp.SUM, group=group)
    per_rank = (n_items + world - 1) // world
    pitch = int(lens.max().item()) if n_items else 0
    pitch = (pitch + 15) // 16 * 16
    mine = torch.zeros(per_rank * max(pitch, 16), dtype=torch.uint8, device=device)
    for slot, i in enumerate(shard_indices(n_items, rank, world)):
        t = local[i]
        mine[slot * pitch: slot * pitch + t.numel()] = t
    allbuf = torch.empty(world * mine.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allbuf, mine, group=group)
    if rank != dst:
        return None
    out = []
    stride = mine.numel()
    for i in range(n_items):
        r, slot = owner_of(i, world), i // world
        base = r * stride + slot * pitch
        out.append(allbuf[base: base + int(lens[i].item())])
    return out


def render_clients_to_grid(acb, client_frames, cfg, grid_w, grid_h, dst=0, group=None):
    """BASELINE config 4 on N GPUs.

    client_frames: dict {client_index: (h,w,3) uint8 torch tensor ON THIS RANK'S GPU} for the clients this rank
                   owns; every client is rendered with the same acb200_render_cfg_t `cfg` (resident batch API),
                   the strings are gathered to `dst` over NCCL and composed there with the device grid kernel.
    Returns (bytes of the grid, n_clients) on dst, None elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    # One explicit (non-default) torch stream carries everything — our kernels, the NCCL collectives and the copies —
    # so their order is the program order.  (A NULL stream handle would mean the library's own internal stream, which
    # does not synchronise with torch's legacy default stream.)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        res = _render_clients_to_grid(acb, client_frames, cfg, grid_w, grid_h, dst, group, world, rank)
    torch.cuda.current_stream().wait_stream(side)
    return res


def _render_clients_to_grid(acb, client_frames, cfg, grid_w, grid_h, dst, group, world, rank):
    n_clients = torch.tensor([len(client_frames)], dtype=torch.int64, device="cuda")
    dist.all_reduce(n_clients, group=group)
    n_clients = int(n_clients.item())
    mine = shard_indices(n_clients, rank, world)
    assert sorted(client_frames) == mine, "clients must be sharded round-robin (client c on rank c % world)"
    local = {}
    if mine:
        batch = torch.stack([client_frames[i] for i in mine]).contiguous()
        n = batch.shape[0]
        cap = acb.frame_capacity(cfg)
        d_out = torch.empty(n * cap, dtype=torch.uint8, device="cuda")
        d_len = torch.empty(n, dtype=torch.int32, device="cuda")
        d_scr = torch.empty(acb.scratch_bytes(cfg, n), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        acb.render_batch_device(cfg, batch.data_ptr(), n, d_out.data_ptr(), cap, d_len.data_ptr(), d_scr.data_ptr(),
                                stream)
        lens = d_len.cpu()
        for slot, i in enumerate(mine):
            local[i] = d_out[slot * cap: slot * cap + int(lens[slot])]
    gathered = gather_variable(local, n_clients, dst=dst, device=torch.device("cuda"), group=group)
    if rank != dst:
        return None
    srcs = [t.contiguous() for t in gathered]
    canvas = torch.empty(max(grid_w * grid_h + grid_h + 1, max(t.numel() for t in srcs) + 1) + 16, dtype=torch.uint8,
                         device="cuda")
    size = acb.create_grid_device([t.data_ptr() for t in srcs], [t.numel() for t in srcs], grid_w, grid_h,
                                  canvas.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().synchronize()
    return canvas[:size].cpu().numpy().tobytes(), n_clients


def arena_slot(item, world, per_rank):
    """index of item's fixed-pitch slot inside the all-gathered arena: rank-major, then the rank's k-th own item"""
    return owner_of(item, world) * per_rank + item // world


def gather_fixed(arena, lens, all_arena, all_lens, group=None):
    """The exchange step of the steady-state grid loop: every rank contributes its fixed-pitch arena (per_rank slots)
    and its per-slot lengths; two fixed-shape all-gathers, no host round trip.  Backend-agnostic (NCCL / gloo)."""
    dist.all_gather_into_tensor(all_arena, arena, group=group)
    dist.all_gather_into_tensor(all_lens, lens, group=group)


class GridPipeline:
    """BASELINE config 4 as a steady-state loop: everything render_clients_to_grid() sizes and allocates per call is
    set up once (fixed-pitch device arenas, gather buffers, pinned host mirrors, one stream), so a grid costs one render
    launch, two fixed-shape collectives, one length read-back, the grid kernel and one copy of the finished canvas.

    The gather payload is the fixed-pitch arena itself (pitch = acb200_frame_capacity): at <= ~120 KB per 160x48
    client string the exchange is latency-bound, so shipping the slack beats a length-dependent second phase.
    """

    def __init__(self, acb, cfg, n_clients, grid_w, grid_h, dst=0, group=None):
        self.acb, self.cfg, self.n, self.W, self.H, self.dst, self.group = acb, cfg, n_clients, grid_w, grid_h, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.mine = shard_indices(n_clients, self.rank, self.world)
        self.per_rank = (n_clients + self.world - 1) // self.world
        self.cap = acb.frame_capacity(cfg)
        dev = torch.device("cuda")
        self.arena = torch.zeros(self.per_rank * self.cap, dtype=torch.uint8, device=dev)
        self.lens = torch.zeros(self.per_rank, dtype=torch.int32, device=dev)
        self.scr = torch.empty(acb.scratch_bytes(cfg, max(1, len(self.mine))), dtype=torch.uint8, device=dev)
        self.all_arena = torch.empty(self.world * self.per_rank * self.cap, dtype=torch.uint8, device=dev)
        self.all_lens = torch.empty(self.world * self.per_rank, dtype=torch.int32, device=dev)
        self.canvas = torch.empty(max(grid_w * grid_h + grid_h + 1, self.cap + 1) + 16, dtype=torch.uint8, device=dev)
        self.h_lens = torch.empty(self.world * self.per_rank, dtype=torch.int32).pin_memory()
        self.h_canvas = torch.empty(self.canvas.numel(), dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream()

    def step(self, batch):
        """batch: contiguous uint8 tensor [len(self.mine), h, w, 3] on this rank's GPU (client self.mine[k] at index k).
        Returns the grid's bytes on dst, None elsewhere."""
        acb, s = self.acb, self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            if self.mine:
                assert batch.is_contiguous() and batch.shape[0] == len(self.mine)
                acb.render_batch_device(self.cfg, batch.data_ptr(), len(self.mine), self.arena.data_ptr(), self.cap,
                                        self.lens.data_ptr(), self.scr.data_ptr(), s.cuda_stream)
            gather_fixed(self.arena, self.lens, self.all_arena, self.all_lens, group=self.group)
            if self.rank != self.dst:
                # the caller may reuse `batch` as soon as this returns: its stream must wait for the render queued here
                torch.cuda.current_stream().wait_stream(s)
                return None
            self.h_lens.copy_(self.all_lens, non_blocking=True)
            s.synchronize()
            base = self.all_arena.data_ptr()
            ptrs, sizes = [], []
            for i in range(self.n):
                k = arena_slot(i, self.world, self.per_rank)
                ptrs.append(base + k * self.cap)
                sizes.append(int(self.h_lens[k]))
            size = acb.create_grid_device(ptrs, sizes, self.W, self.H, self.canvas.data_ptr(), s.cuda_stream)
            self.h_canvas[:size].copy_(self.canvas[:size], non_blocking=True)
            s.synchronize()
            return self.h_canvas[:size].numpy().tobytes()


class PixelGridPipeline:
    """BASELINE config 4, pixel-space (the server's compositor, src/server/stream.c:664-779), sharded over the ranks:
    SURVEY.md §8e "gather the NN-resized cell images".  Rank r holds the frames of clients c % world == r, resizes
    each to the size create_multi_source_composite would give it in a W x 2H composite (acb200_mixed_cell_size +
    acb200_resize_nn_device: a few KB per client instead of 6.2 MB), one fixed-pitch all-gather moves the cell images,
    and the composing rank blits + converts them with acb200_mixed_frame_device(prefit=1): byte-identical to
    create_mixed_ascii_frame_for_client on the full frames."""

    def __init__(self, acb, sizes, W, H, caps, palette, dst=0, group=None):
        """sizes: [(w, h)] of ALL n clients (every rank knows the geometry; only pixels are sharded)"""
        self.acb, self.W, self.H, self.caps, self.palette, self.dst, self.group = acb, W, H, caps, palette, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = len(sizes)
        self.ws, self.hs = [s[0] for s in sizes], [s[1] for s in sizes]
        self.mine = shard_indices(self.n, self.rank, self.world)
        self.per_rank = (self.n + self.world - 1) // self.world
        self.cell = [acb.mixed_cell_size(self.ws, self.hs, i, W, H) for i in range(self.n)]
        self.pitch = (max(tw * th * 3 for tw, th in self.cell) + 255) // 256 * 256 or 256
        dev = torch.device("cuda")
        self.local = torch.zeros(self.per_rank * self.pitch, dtype=torch.uint8, device=dev)
        self.all = torch.empty(self.world * self.per_rank * self.pitch, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.Stream()

    def step(self, frames):
        """frames: list of contiguous uint8 CUDA tensors (h, w, 3), frames[k] = client self.mine[k].
        Returns the frame string on dst, None elsewhere."""
        acb, s = self.acb, self.stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for k, i in enumerate(self.mine):
                tw, th = self.cell[i]
                if tw > 0 and th > 0 and self.n > 1:
                    acb.resize_nn_device(frames[k].data_ptr(), self.ws[i], self.hs[i],
                                         self.local.data_ptr() + k * self.pitch, tw, th, s.cuda_stream)
                elif self.n == 1:
                    self.local[: frames[k].numel()] = frames[k].reshape(-1)
            dist.all_gather_into_tensor(self.all, self.local, group=self.group)
            if self.rank != self.dst:
                torch.cuda.current_stream().wait_stream(s)
                return None
            s.synchronize()  # acb200_mixed_frame_device runs on the library's own stream for this thread
            base = self.all.data_ptr()
            ptrs = [base + arena_slot(i, self.world, self.per_rank) * self.pitch for i in range(self.n)]
            return acb.mixed_frame_device(ptrs, self.ws, self.hs, True, self.W, self.H, self.caps, self.palette)
