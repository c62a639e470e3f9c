"""world_size-2 (and 3) gloo tests of the N>1 host logic: round-robin client sharding and the variable-length
gather that feeds the grid compositor (ascii-chat_b200/multi.py).  No GPU: the per-client "render" is stood in
by the oracle port, which is exactly what the GPU path is proven byte-equal to in test_gpu_parity.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clients, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("acb_multi", os.path.join(ROOT, "ascii-chat_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    import oracle_bind as ob

    mine = multi.shard_indices(n_clients, rank, world)
    local = {}
    for c in mine:
        img = ob.gen(("noise", "bars", "gradient")[c % 3], 96, 64, c)
        s = ob.port_convert(img, 20 + c, 6 + (c % 3), 3 if c % 2 else 0, 2 if c % 2 else 0)  # ragged lengths
        local[c] = torch.frombuffer(bytearray(s), dtype=torch.uint8)
    got = multi.gather_variable(local, n_clients, dst=0, device=torch.device("cpu"))
    if rank == 0:
        frames = [bytes(t.numpy().tobytes()) for t in got]
        grid, size = ob.port_create_grid(frames, 100, 30)
        q.put((frames, grid, size))
    dist.barrier()
    dist.destroy_process_group()


def _worker_fixed(rank, world, port, n_clients, q):
    """the steady-state exchange of multi.GridPipeline (fixed-pitch arena + lengths, two all-gathers) over gloo"""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("acb_multi", os.path.join(ROOT, "ascii-chat_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    import oracle_bind as ob

    cap = 4096
    per_rank = (n_clients + world - 1) // world
    arena = torch.zeros(per_rank * cap, dtype=torch.uint8)
    lens = torch.zeros(per_rank, dtype=torch.int32)
    for k, c in enumerate(multi.shard_indices(n_clients, rank, world)):
        s = ob.port_convert(ob.gen(("noise", "bars", "gradient")[c % 3], 96, 64, c), 10 + c, 4 + (c % 3), 2, 0)
        assert len(s) <= cap
        arena[k * cap: k * cap + len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8)
        lens[k] = len(s)
    all_arena = torch.empty(world * per_rank * cap, dtype=torch.uint8)
    all_lens = torch.empty(world * per_rank, dtype=torch.int32)
    for _ in range(2):  # the buffers are reused every step
        multi.gather_fixed(arena, lens, all_arena, all_lens)
    if rank == 0:
        frames = []
        for i in range(n_clients):
            k = multi.arena_slot(i, world, per_rank)
            frames.append(all_arena[k * cap: k * cap + int(all_lens[k])].numpy().tobytes())
        q.put(frames)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clients", [(2, 8), (2, 5), (3, 7)])
def test_fixed_pitch_gather_addressing(world, n_clients):
    sys.path[:0] = [os.path.join(ROOT, "tests")]
    import oracle_bind as ob
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fixed, args=(r, world, port, n_clients, q)) for r in range(world)]
    [p.start() for p in procs]
    frames = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    exp = [ob.port_convert(ob.gen(("noise", "bars", "gradient")[c % 3], 96, 64, c), 10 + c, 4 + (c % 3), 2, 0)
           for c in range(n_clients)]
    assert frames == exp


@pytest.mark.parametrize("world,n_clients", [(2, 8), (2, 5), (3, 7), (2, 1)])
def test_sharded_gather_equals_single_process(world, n_clients):
    sys.path[:0] = [os.path.join(ROOT, "tests")]
    import oracle_bind as ob
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clients, q)) for r in range(world)]
    [p.start() for p in procs]
    frames, grid, size = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    exp_frames = []
    for c in range(n_clients):
        img = ob.gen(("noise", "bars", "gradient")[c % 3], 96, 64, c)
        exp_frames.append(ob.port_convert(img, 20 + c, 6 + (c % 3), 3 if c % 2 else 0, 2 if c % 2 else 0))
    assert frames == exp_frames
    assert (grid, size) == ob.port_create_grid(exp_frames, 100, 30)


def test_shard_indices_partition():
    sys.path[:0] = [ROOT]
    import importlib.util
    spec = importlib.util.spec_from_file_location("acb_multi", os.path.join(ROOT, "ascii-chat_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 9, 100):
            parts = [multi.shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert all(multi.owner_of(i, world) == r for r, p in enumerate(parts) for i in p)
