"""BASELINE config 4 measured: 8 client streams of 1920x1080, sharded round-robin over the ranks, each rendered to
160x48 ANSI-256 with the resident batch API, strings gathered over NCCL, text grid composed on rank 0
(multi.render_clients_to_grid).  Run under torchrun; prints one JSON line on rank 0.  The path is latency-bound
(<= 90 KB per client string): the number that matters is grids per second, max over ranks."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import ascii_chat_b200 as acb
    from ascii_chat_b200 import multi
    import oracle_bind as ob
    assert acb.lib().acb200_init(lr) == 0
    n_clients, W, H, cols, rows, level, mode = 8, 1920, 1080, 160, 48, 2, 0
    out = {}
    for scale, name in ((acb.SCALE_NN, "nn"), (acb.SCALE_BOX, "box")):
        cfg = acb.make_cfg(W, H, cols, rows, level, mode, scale=scale)
        mine = {c: torch.from_numpy(ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c)).cuda()
                for c in multi.shard_indices(n_clients, rank, world)}
        res = None
        for _ in range(20):
            res = multi.render_clients_to_grid(acb, mine, cfg, 320, 96)
        torch.cuda.synchronize()
        dist.barrier()
        K = 100
        t0 = time.perf_counter()
        for _ in range(K):
            res = multi.render_clients_to_grid(acb, mine, cfg, 320, 96)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the steady-state form: buffers and streams set up once (multi.GridPipeline)
        pipe = multi.GridPipeline(acb, cfg, n_clients, 320, 96)
        batch = torch.stack([mine[c] for c in pipe.mine]).contiguous() if pipe.mine else None
        res2 = None
        for _ in range(20):
            res2 = pipe.step(batch)
        torch.cuda.synchronize()
        dist.barrier()
        K2 = 300
        t0 = time.perf_counter()
        for _ in range(K2):
            res2 = pipe.step(batch)
        torch.cuda.synchronize()
        dt2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
        if rank == 0:
            frames = [ob.port_convert(ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c), cols, rows, level, mode,
                                      scale=ob.SCALE_BOX if scale == acb.SCALE_BOX else ob.SCALE_NN)
                      for c in range(n_clients)]
            exp, size = ob.port_create_grid(frames, 320, 96)
            out[name] = {"grids_per_s": K / float(dt), "ms_per_grid": 1e3 * float(dt) / K,
                         "source_Mpix_s": K * n_clients * W * H / 1e6 / float(dt),
                         "bytes_identical_to_oracle": bool(res[0] == exp or res[0] == exp[:size]), "grid_bytes": len(res[0]),
                         "pipeline_grids_per_s": K2 / float(dt2), "pipeline_ms_per_grid": 1e3 * float(dt2) / K2,
                         "pipeline_source_Mpix_s": K2 * n_clients * W * H / 1e6 / float(dt2),
                         "pipeline_bytes_identical_to_oracle": bool(res2 == exp or res2 == exp[:size])}
    if rank == 0:
        print(json.dumps({"config": "C4: 8 clients x 1920x1080 -> 160x48 ANSI-256, text grid 320x96 on rank 0",
                          "n_gpus": world, "collective": "NCCL all_reduce(lengths) + all_gather(fixed pitch)", **out}))
    dist.barrier()
    dist.destroy_process_group()


main()
