"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards two ways:
  * scenes are independent           -> scene-parallel, no data-path collective (what bench.py --gpus N measures);
  * frames of one scene are independent up to the per-cell winner rule -> each rank fuses its frame shard into its
    own store, then ONE all-gather of the compacted per-rank deltas {cell, key, h, token row} and every rank applies
    the same winner rule (bscnav_merge_delta).  The rule is a min over (float32 r^2, ordinal) -- commutative and
    associative -- so the merged store does not depend on the number of ranks or on arrival order.
  * queries over a grid whose cells are spread over ranks (tile-sharded scene): local top-K per rank, ONE all-gather of K x (similarity, position) per query, K-way merge on every rank
    (`query_sharded`, `merge_topk`) -- 1.6 KB per query and rank at K = 100.
The reference has none of this (CUDA_VISIBLE_DEVICES='0', create_memory_for_dataset.py:38).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous frame shard of a scene for `rank` (sizes differ by at most one; ordinals stay global)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_scenes(n_scenes: int, rank: int, world: int) -> list:
    """Round-robin scene assignment (scene-parallel build / query replicas)."""
    return list(range(rank, n_scenes, world))


def gather_deltas(d_cell: torch.Tensor, d_key: torch.Tensor, d_h: torch.Tensor, d_feat: torch.Tensor, n: int, group=None):
    """All-gather variable-length deltas.  Inputs hold `n` valid entries (rows beyond n are ignored).
    Protocol: all-gather the counts, pad every payload to the maximum count, one all-gather per array, then strip the
    padding in rank order.  Works on any device/backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    dev = d_cell.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.cpu().tolist()
    cap = max(max(counts), 1)

    def padded(t):
        out = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        out[:n] = t[:n]
        return out

    outs = []
    for t in (d_cell, d_key, d_h, d_feat):
        buf = torch.empty((world * cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(buf, padded(t), group=group)
        parts = [buf[r * cap: r * cap + counts[r]] for r in range(world)]
        outs.append(torch.cat(parts) if parts else buf[:0])
    return tuple(outs) + (counts,)


def merge_scene_across_ranks(store, ord_lo: int, ord_hi: int, max_rows: int, group=None):
    """Export the rows this rank produced (ordinals in [ord_lo, ord_hi)), all-gather them, and merge every other rank's
    rows into the local store.  After the call all ranks hold the same kept-token set for the scene."""
    rank = dist.get_rank(group)
    d_cell, d_key, d_h, d_feat, n_dev = store.export_delta(ord_lo, ord_hi, max_rows)
    n = int(n_dev.item())
    g_cell, g_key, g_h, g_feat, counts = gather_deltas(d_cell, d_key, d_h, d_feat, n, group)
    off = 0
    for r, c in enumerate(counts):
        if r != rank and c:
            store.merge_delta(g_cell[off:off + c].contiguous(), g_key[off:off + c].contiguous(), g_h[off:off + c].contiguous(),
                              g_feat[off:off + c].contiguous())
        off += c
    return counts


def merge_rule_host(cells: np.ndarray, keys: np.ndarray, cache: int) -> dict:
    """Host statement of the winner rule used by bscnav_merge_delta (tests): per cell the `cache` smallest uint64 keys."""
    out = {}
    for c, k in zip(cells.tolist(), keys.tolist()):
        if c >= 0:
            out.setdefault(c, []).append(k)
    return {c: sorted(v)[:cache] for c, v in out.items()}


def merge_topk(sims: torch.Tensor, pos: torch.Tensor, counts: torch.Tensor, k: int):
    """K-way merge of per-rank query results for ONE query batch: sims (R, Q, K) f32, pos (R, Q, K, 3) i32, counts (R, Q) i32
    -> (sims (Q, k), pos (Q, k, 3), counts (Q,)) with the layout of SpatialStore.query (entries past the count: -inf / -1).
    A cell reported by several ranks keeps its largest similarity (the reference's per-voxel max, memory_2.py:655-662);
    order: similarity descending, ties by (row, col, h) ascending -- independent of the rank count and order."""
    R, Q, K = sims.shape
    dev = sims.device
    s = sims.permute(1, 0, 2).reshape(Q, R * K)
    p = pos.permute(1, 0, 2, 3).reshape(Q, R * K, 3).long()
    valid = (torch.arange(K, device=dev)[None, None, :] < counts.permute(1, 0)[:, :, None].long()).reshape(Q, R * K)
    cell = (p[..., 0] << 42) | (p[..., 1] << 21) | p[..., 2]               # grid coordinates are < 2^21
    cell = torch.where(valid, cell, torch.full_like(cell, torch.iinfo(torch.int64).max))
    s = torch.where(valid, s, torch.full_like(s, float("-inf")))
    # order by (similarity desc, cell asc): stable sort by the minor key first
    o1 = torch.argsort(cell, dim=1, stable=True)
    s, cell, p, valid = s.gather(1, o1), cell.gather(1, o1), p.gather(1, o1[..., None].expand(-1, -1, 3)), valid.gather(1, o1)
    o2 = torch.argsort(s, dim=1, descending=True, stable=True)
    s, cell, p, valid = s.gather(1, o2), cell.gather(1, o2), p.gather(1, o2[..., None].expand(-1, -1, 3)), valid.gather(1, o2)
    # a cell's first occurrence in this order carries its maximum: drop the later ones
    n = R * K
    dup = (cell[:, :, None] == cell[:, None, :]) & torch.ones(n, n, dtype=torch.bool, device=dev).tril(-1)[None]
    keep = valid & ~dup.any(dim=2)
    rank = torch.cumsum(keep, dim=1) - 1
    out_s = torch.full((Q, k), float("-inf"), dtype=sims.dtype, device=dev)
    out_p = torch.full((Q, k, 3), -1, dtype=pos.dtype, device=dev)
    take = keep & (rank < k)
    qi = torch.arange(Q, device=dev)[:, None].expand(-1, n)[take]
    out_s[qi, rank[take]] = s[take]
    out_p[qi, rank[take]] = p[take].to(pos.dtype)
    return out_s, out_p, torch.clamp(keep.sum(dim=1), max=k).to(counts.dtype)


def query_sharded(store, q: torch.Tensor, k: int = 100, group=None, batch: bool = False, **filters):
    """Query a scene whose cells are spread over the ranks of `group`: every rank scores its own store (exact path, or the
    batch path with batch=True), the K best of each rank are all-gathered and merged.  Every rank returns the same
    (sims (Q, k), positions (Q, k, 3), counts (Q,)) as a single store holding all the cells would."""
    sims, pos, cnt = (store.query_batch if batch else store.query)(q, k=k, **filters)
    world = dist.get_world_size(group)
    outs = []
    for t in (sims, pos, cnt):                       # concatenated along dim 0 (the layout every backend accepts), then (R, ...)
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        outs.append(buf.reshape((world,) + tuple(t.shape)))
    g_s, g_p, g_c = outs
    return merge_topk(g_s, g_p, g_c, k)
