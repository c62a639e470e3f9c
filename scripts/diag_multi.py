"""2-rank NCCL grid diagnosis: prints where the gathered frames / the grid first differ from the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, torch.distributed as dist

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import ascii_chat_b200 as acb
    from ascii_chat_b200 import multi
    import oracle_bind as ob
    assert acb.lib().acb200_init(rank) == 0
    for n_clients, level, mode in ((8, 0, 0), (8, 3, 2), (5, 2, 0)):
        W, H, cols, rows = 480, 270, 80, 24
        cfg = acb.make_cfg(W, H, cols, rows * 2 if mode == 2 else rows, level, mode)
        mine = {c: torch.from_numpy(ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c)).cuda()
                for c in multi.shard_indices(n_clients, rank, world)}
        res = multi.render_clients_to_grid(acb, mine, cfg, 160, 48)
        if rank == 0:
            frames = [ob.port_convert(ob.gen(("noise", "bars", "gradient")[c % 3], W, H, c), cols, rows, level, mode)
                      for c in range(n_clients)]
            exp, size = ob.port_create_grid(frames, 160, 48)
            grid, _ = res
            ok = grid == exp
            k = next((i for i in range(min(len(grid), len(exp))) if grid[i] != exp[i]), -1)
            print("case", n_clients, level, mode, "grid ok" if ok else "GRID MISMATCH", len(grid), len(exp), size, "first diff", k,
                  (grid[max(0, k - 20):k + 30], exp[max(0, k - 20):k + 30]) if not ok else "")
    dist.barrier(); dist.destroy_process_group()
main()
