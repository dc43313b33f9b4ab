"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards two ways:
  * scenes are independent           -> scene-parallel, no data-path collective (what bench.py --gpus N measures);
  * frames of one scene are independent up to the per-cell winner rule -> each rank fuses its frame shard into its
    own store, then ONE all-gather of the compacted per-rank deltas {cell, key, h, token row} and every rank applies
    the same winner rule (bscnav_merge_delta).  The rule is a min over (float32 r^2, ordinal) -- commutative and
    associative -- so the merged store does not depend on the number of ranks or on arrival order.
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
