"""torchrun worker (2 ranks, NCCL): frames of one scene sharded over ranks -> export / all-gather / merge -> every rank's
store equals the single-process joint build.  Launched by tests/test_gpu_merge.py::test_two_rank_nccl_merge."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("bsc-nav_b200")
    dmod = importlib.import_module("bsc-nav_b200.dist")
    H, W, GS, D = 96, 128, 64, 128
    counts = dmod.merge_self_check()                                   # sharded build + packed all-gather merge == joint build
    for cache, bev in ((1, True), (10, False)):
        dmod.merge_self_check(cache=cache, bev=bev)
    # queries over a scene whose grid rows are spread over the ranks: local top-K + one all-gather + K-way merge
    # equals the query over one store holding every cell
    qcfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=2, bev=False,
                           max_slots=4096)
    gen = torch.Generator("cuda").manual_seed(3)                      # same stream of numbers on every rank
    whole, part = pkg.SpatialStore(qcfg), pkg.SpatialStore(qcfg)
    levels, n_cells, C = whole.g.levels, 2500, 2
    cells = torch.randperm(GS * GS * levels, device="cuda", generator=gen)[:n_cells].to(torch.int32)
    base = torch.randn(150, D, device="cuda", generator=gen)
    feat = base[torch.randint(0, 150, (n_cells * C,), device="cuda", generator=gen)]           # repeated rows: ties
    tok_h = torch.randint(0, 30, (n_cells * C,), device="cuda", generator=gen).to(torch.int32)

    def install(s, sel):
        rows = (sel[:, None] * C + torch.arange(C, device="cuda")[None]).flatten()
        m = len(sel)
        s.reset()
        s.feat[:m * C] = feat[rows]
        s.slot_cell[:m] = cells[sel]
        s.occupied_ids[cells[sel].long()] = torch.arange(m, dtype=torch.int32, device="cuda")
        s.row_key[:m * C] = 0
        s.slot_keys[:m * C] = 0
        s.tok_h[:m * C] = tok_h[rows]
        s.row_norms(m * C)
        s.counters[0] = m

    install(whole, torch.arange(n_cells, device="cuda"))
    grid_row = (cells.long() // levels) // GS
    install(part, (grid_row % world == rank).nonzero().flatten())
    q = torch.cat([base[:3] + 0.05 * torch.randn(3, D, device="cuda", generator=gen), torch.randn(2, D, device="cuda", generator=gen)])
    for use_batch in (False, True):
        got = dmod.query_sharded(part, q, k=50, batch=use_batch)
        want = whole.query(q, k=50)
        for a, b, name in zip(got, want, ("sims", "positions", "counts")):
            assert torch.equal(a, b), f"rank {rank}: sharded query {name} differ from the whole store (batch={use_batch})"
    dist.barrier()
    if rank == 0:
        print("MERGE_OK", counts)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
