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
    rows into the local store.  After the call all ranks hold the same kept-token set for the scene.

    One packed buffer per rank ({cell, h, key, token row}, SpatialStore.export_packed), ONE all-gather of it over NVLink and
    one merge launch over the gathered rows (the per-rank counts travel as a second, 8-byte all-gather and are read by the
    merge kernels on the device).  The only host wait is the read of the maximum count, which sizes the gathered buffer
    (all-gather needs equal shares; padding to the store's capacity would move a gigabyte per rank)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = store.device
    packed, n_dev = store.export_packed(ord_lo, ord_hi, max_rows)
    counts = torch.empty(world * 2, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(counts, n_dev, group=group)
    host = counts.view(world, 2).cpu()                       # the one host round trip: sizes the collective
    if bool((host[:, 1] > max_rows).any()):
        raise RuntimeError(f"export truncated: a rank produced {int(host[:, 1].max())} rows > max_rows = {max_rows}")
    cap = max(int(host[:, 0].max()), 1)
    gathered = torch.empty(world * cap, packed.shape[1], dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, packed[:cap], group=group)
    store.merge_packed(gathered, world, cap, counts, count_stride=2, skip_segment=rank)
    return host[:, 0].tolist()


def fuse_scene_sharded(mem, poses, load_frames, batch: int, group=None, max_rows=None):
    """Build ONE scene with its frames sharded over the ranks of `group` (BASELINE configs[4]): rank r fuses the contiguous shard
    shard_frames(F, r, world) through the facade's batched build step (VoxelTokenMemory.fuse_frames: host buffers in, copies and
    scatter overlapped with the ViT), then the ranks exchange what they stored (merge_scene_across_ranks: one packed all-gather)
    so that every rank ends with the kept-token set of the whole scene.  poses: (F, 7) of the whole scene, the same on every
    rank; load_frames(f0, n) -> (rgb (n,H,W,3|4) uint8, depth (n,H,W) float32) for frames f0 .. f0 + n - 1.  Returns the per-rank
    export counts (None on one rank)."""
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 7)
    shard = shard_frames(len(poses), rank, world)
    mem.begin_shard(poses[0], shard.start)
    for f0 in range(shard.start, shard.stop, batch):
        n = min(batch, shard.stop - f0)
        rgb, depth = load_frames(f0, n)
        mem.fuse_frames(rgb, depth, poses[f0:f0 + n])
    mem._join()
    if world == 1:
        return None
    npts = mem.args.height * mem.args.width
    return merge_scene_across_ranks(mem._store, shard.start * npts, shard.stop * npts,
                                    max_rows if max_rows is not None else mem._store.g.max_slots * mem._store.g.cache, group)


def merge_self_check(group=None, cache: int = 3, bev: bool = False) -> list:
    """Parity check of the collective path on the ranks of `group` (run by tests/multi_rank_merge.py and once by every
    `bench.py --gpus N` run, N > 1): a small scene whose frames are sharded over the ranks (shard_frames), fused per rank and
    merged with merge_scene_across_ranks, must leave EVERY rank with exactly the store a single-process joint build produces
    (cell set, kept keys, heights and token rows bit for bit).  Raises on any difference; returns the per-rank export counts."""
    from . import synth
    from .geometry import PoseChain
    from .store import SpatialStore, StoreConfig
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    H, W, GS, D = 96, 128, 64, 128
    n = 4 * world + 1
    chain = PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    frames = [synth.synthetic_frame(f, H, W, n_frames=n + 3) for f in range(n)]
    depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    poses = np.stack([chain(f[2]) for f in frames])
    tokens = torch.randn(n, 256, D, device="cuda", generator=torch.Generator("cuda").manual_seed(1))   # same on every rank
    cfg = StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=cache, bev=bev,
                      max_slots=16384)
    ob = np.arange(n, dtype=np.int64) * H * W
    joint = SpatialStore(cfg)
    joint.fuse(depth, poses, tokens, ord_base=ob)
    mine = list(shard_frames(n, rank, world))
    st = SpatialStore(cfg)
    st.fuse(depth[mine].contiguous(), poses[mine], tokens[mine].contiguous(), ord_base=ob[mine])
    counts = merge_scene_across_ranks(st, int(ob[mine[0]]), int(ob[mine[-1]]) + H * W, st.g.max_slots * cache, group=group)

    def canon(s):
        return {k: [(float(t[0]), int(t[1]), int(t[2]), t[3].tobytes()) for t in v] for k, v in s.snapshot().items()}
    got, want = canon(st), canon(joint)
    if got != want:
        bad = [k for k in set(got) | set(want) if got.get(k) != want.get(k)]
        raise RuntimeError(f"rank {rank}: merged store differs from the joint build in {len(bad)} cells (first {sorted(bad)[:3]})")
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
    # group by cell with the best similarity first: stable sort by similarity (desc), then by cell (asc)
    def take(o, *ts):
        return [t.gather(1, o[..., None].expand(-1, -1, 3)) if t.dim() == 3 else t.gather(1, o) for t in ts]
    s, cell, p, valid = take(torch.argsort(s, dim=1, descending=True, stable=True), s, cell, p, valid)
    s, cell, p, valid = take(torch.argsort(cell, dim=1, stable=True), s, cell, p, valid)
    # a cell's first occurrence carries its maximum: later ones are dropped (neighbour comparison, O(R K) memory per query)
    first = torch.ones_like(valid)
    first[:, 1:] = cell[:, 1:] != cell[:, :-1]
    keep = valid & first
    # final order: similarity desc, ties by cell asc (the entries are in cell order, the sort is stable)
    s = torch.where(keep, s, torch.full_like(s, float("-inf")))
    s, p, keep = take(torch.argsort(s, dim=1, descending=True, stable=True), s, p, keep)
    n = R * K
    if n < k:
        s = torch.cat([s, torch.full((Q, k - n), float("-inf"), dtype=s.dtype, device=dev)], dim=1)
        p = torch.cat([p, torch.full((Q, k - n, 3), -1, dtype=p.dtype, device=dev)], dim=1)
        keep = torch.cat([keep, torch.zeros(Q, k - n, dtype=torch.bool, device=dev)], dim=1)
    out_s = s[:, :k].contiguous()
    out_p = torch.where(keep[:, :k, None], p[:, :k], torch.full_like(p[:, :k], -1)).to(pos.dtype)
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
