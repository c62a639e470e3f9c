"""Multi-GPU grid path (BASELINE config 4) on real GPUs: 8 client streams sharded round-robin over the ranks,
rendered with the resident batch API, gathered over NCCL, composed on rank 0 with the device grid kernel.
The result must equal the single-process oracle composition byte for byte.  Skipped on a 1-GPU box
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C4 = (1920, 1080, 160, 48)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clients, level, mode, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ascii_chat_b200 as acb
    from ascii_chat_b200 import multi
    import oracle_bind as ob
    assert acb.lib().acb200_init(rank) == 0
    W, H, cols, rows = C4  # BASELINE config 4: 1080p clients -> 160x48 cells -> 320x96 grid
    cfg = acb.make_cfg(W, H, cols, rows * 2 if mode == 2 else rows, level, mode)
    mine = {c: torch.from_numpy(ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c)).cuda()
            for c in multi.shard_indices(n_clients, rank, world)}
    res = multi.render_clients_to_grid(acb, mine, cfg, 320, 96)
    # the steady-state pipeline (buffers set up once) must produce the same grid, twice in a row
    pipe = multi.GridPipeline(acb, cfg, n_clients, 320, 96)
    batch = torch.stack([mine[c] for c in pipe.mine]).contiguous() if pipe.mine else None
    for _ in range(2):
        res2 = pipe.step(batch)
        if rank == 0:
            assert res2 == res[0], "GridPipeline differs from render_clients_to_grid"
    # pixel-space: gather the NN-resized cell images, composite + convert on rank 0 (the server's compositor)
    pix = []
    for (vw, vh) in ((160, 48), (320, 96)):
        pp = multi.PixelGridPipeline(acb, [(W, H)] * n_clients, vw, vh, acb.make_caps(level, mode, True), "standard")
        frames = [mine[c] for c in pp.mine]
        for _ in range(2):
            out = pp.step(frames)
        pix.append(out)
    if rank == 0:
        q.put((res, pix))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clients,level,mode", [(8, 0, 0), (8, 3, 2), (5, 2, 0)])
def test_grid_over_nccl(n_clients, level, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    sys.path[:0] = [os.path.join(ROOT, "tests")]
    import oracle_bind as ob
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clients, level, mode, q)) for r in range(world)]
    [p.start() for p in procs]
    (grid, n), pix = q.get(timeout=300)
    [p.join(timeout=120) for p in procs]
    assert n == n_clients and all(p.exitcode == 0 for p in procs)
    # checker: the COMPILED reference at BASELINE config 4's full size (the port only if oracle/_ref did not travel)
    W, H, cols, rows = C4
    conv = ob.ref_convert if ob.ref() is not None else ob.port_convert
    mk_grid = ob.ref_create_grid if ob.ref() is not None else ob.port_create_grid
    mixed = ob.ref_mixed_frame if ob.ref() is not None else ob.port_mixed_frame
    srcs = [ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c) for c in range(n_clients)]
    exp, size = mk_grid([conv(s, cols, rows, level, mode) for s in srcs], 320, 96)
    assert grid == exp[:size] or grid == exp
    for got, (vw, vh) in zip(pix, ((160, 48), (320, 96))):
        assert got == mixed(srcs, vw, vh, level, mode, "standard", True)[0], (vw, vh)
