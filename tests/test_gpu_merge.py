"""GPU: the multi-GPU merge path (export_delta -> all-gather -> merge_delta).  Single-GPU tests simulate the rank
shards with separate stores; the 2-rank NCCL test runs when the box has two GPUs."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import memory_port as mp

pytestmark = pytest.mark.gpu

H, W, GS, D = 96, 128, 64, 128
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(pkg, n_frames):
    chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    frames = [mp.synthetic_frame(f, H, W, n_frames=12) for f in range(n_frames)]
    depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    poses = np.stack([chain(f[2]) for f in frames])
    tokens = torch.randn(n_frames, 256, D, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    return depth, poses, tokens


def _canon(st):
    """Order-independent view of a store: {cell key: [(r2, ordinal, h, checksum of the feature row)]}."""
    return {k: [(float(t[0]), int(t[1]), int(t[2]), float(np.dot(t[3].astype(np.float64), np.arange(1, t[3].size + 1))))
                for t in v] for k, v in st.snapshot().items()}


def _diff(a, b):
    out = [f"{k}: {a.get(k)} != {b.get(k)}" for k in sorted(set(a) | set(b)) if a.get(k) != b.get(k)]
    return f"{len(out)} differing cells; first: " + " | ".join(out[:3])


@pytest.mark.parametrize("cache,bev", [(1, True), (4, False), (10, False)])
def test_merge_equals_joint_build(pkg, cache, bev):
    n = 8
    depth, poses, tokens = _inputs(pkg, n)
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=cache, bev=bev)
    ob = np.arange(n, dtype=np.int64) * H * W
    joint = pkg.SpatialStore(cfg)
    joint.fuse(depth, poses, tokens, ord_base=ob)
    shards = [range(0, 3), range(3, 8)]                  # uneven "ranks"
    stores = []
    for sh in shards:
        s = pkg.SpatialStore(cfg)
        idx = list(sh)
        s.fuse(depth[idx].contiguous(), poses[idx], tokens[idx].contiguous(), ord_base=ob[idx])
        stores.append(s)
    deltas = []
    for s, sh in zip(stores, shards):
        d_cell, d_key, d_h, d_feat, n_dev = s.export_delta(int(ob[sh[0]]), int(ob[sh[-1]]) + H * W, max_out=s.g.max_slots * cache)
        k = int(n_dev[0].item())
        assert k == sum(len(v) for v in s.snapshot().values())
        deltas.append((d_cell[:k].contiguous(), d_key[:k].contiguous(), d_h[:k].contiguous(), d_feat[:k].contiguous()))
    stores[0].merge_delta(*deltas[1])
    stores[1].merge_delta(*deltas[0])
    want = _canon(joint)
    got0, got1 = _canon(stores[0]), _canon(stores[1])
    assert got0 == want, _diff(got0, want)
    assert got1 == want, _diff(got1, want)
    # merging a delta twice changes nothing (idempotent)
    stores[0].merge_delta(*deltas[1])
    got0 = _canon(stores[0])
    assert got0 == want, _diff(got0, want)


@pytest.mark.parametrize("cache,bev,n_shards", [(1, True, 2), (3, False, 3), (1, False, 4)])
def test_packed_merge_equals_joint_build(pkg, cache, bev, n_shards):
    """The all-gather form: every shard exports ONE packed buffer, the buffers are laid side by side padded to the largest
    count (what all_gather_into_tensor produces), one merge_packed launch per store with its own segment skipped."""
    n = 9
    depth, poses, tokens = _inputs(pkg, n)
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=cache, bev=bev,
                          max_slots=16384)
    ob = np.arange(n, dtype=np.int64) * H * W
    joint = pkg.SpatialStore(cfg)
    joint.fuse(depth, poses, tokens, ord_base=ob)
    dmod = importlib.import_module("bsc-nav_b200.dist")
    shards = [list(dmod.shard_frames(n, r, n_shards)) for r in range(n_shards)]
    stores, packed, counts = [], [], []
    for idx in shards:
        s = pkg.SpatialStore(cfg)
        s.fuse(depth[idx].contiguous(), poses[idx], tokens[idx].contiguous(), ord_base=ob[idx])
        buf, n_dev = s.export_packed(int(ob[idx[0]]), int(ob[idx[-1]]) + H * W, max_out=s.g.max_slots * cache)
        assert int(n_dev[0]) == int(n_dev[1]) == sum(len(v) for v in s.snapshot().values())
        stores.append(s); packed.append(buf); counts.append(n_dev)
    cap = max(int(c[0]) for c in counts)
    gathered = torch.cat([b[:cap] for b in packed]).contiguous()
    seg_counts = torch.cat(counts).contiguous()
    want = _canon(joint)
    for r, s in enumerate(stores):
        s.merge_packed(gathered, n_shards, cap, seg_counts, count_stride=2, skip_segment=r)
        got = _canon(s)
        assert got == want, _diff(got, want)
    stores[0].merge_packed(gathered, n_shards, cap, seg_counts, count_stride=2, skip_segment=-1)     # idempotent, own rows included
    assert _canon(stores[0]) == want


def test_export_reports_truncation(pkg):
    depth, poses, tokens = _inputs(pkg, 2)
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=1, bev=True)
    s = pkg.SpatialStore(cfg)
    s.fuse(depth, poses, tokens)
    total = s.count()
    _, n_dev = s.export_packed(0, 2 ** 32 - 1, max_out=10)
    assert n_dev.tolist() == [10, total]
    *_, n_dev = s.export_delta(0, 2 ** 32 - 1, max_out=10)
    assert n_dev.tolist() == [10, total]


def test_query_after_merge_matches_joint(pkg):
    n = 6
    depth, poses, tokens = _inputs(pkg, n)
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=2, bev=False)
    ob = np.arange(n, dtype=np.int64) * H * W
    joint = pkg.SpatialStore(cfg)
    joint.fuse(depth, poses, tokens, ord_base=ob)
    a, b = pkg.SpatialStore(cfg), pkg.SpatialStore(cfg)
    a.fuse(depth[:3].contiguous(), poses[:3], tokens[:3].contiguous(), ord_base=ob[:3])
    b.fuse(depth[3:].contiguous(), poses[3:], tokens[3:].contiguous(), ord_base=ob[3:])
    d_cell, d_key, d_h, d_feat, n_dev = b.export_delta(int(ob[3]), int(ob[-1]) + H * W, max_out=b.g.max_slots * 2)
    k = int(n_dev[0].item())
    a.merge_delta(d_cell[:k].contiguous(), d_key[:k].contiguous(), d_h[:k].contiguous(), d_feat[:k].contiguous())
    q = tokens[2, 77:79].contiguous()
    s1, p1, c1 = joint.query(q, k=50)
    s2, p2, c2 = a.query(q, k=50)
    assert torch.equal(c1, c2) and torch.equal(p1, p2) and torch.equal(s1, s2)


def _install_cells(st, cells, feat, tok_h):
    """Fill a store by hand: slot i holds cell cells[i] with the C token rows feat[i*C:(i+1)*C]."""
    n, C = len(cells), st.g.cache
    st.reset()
    st.feat[:n * C] = feat
    st.slot_cell[:n] = cells
    st.occupied_ids[cells.long()] = torch.arange(n, dtype=torch.int32, device="cuda")
    st.row_key[:n * C] = 0
    st.slot_keys[:n * C] = 0
    st.tok_h[:n * C] = tok_h
    st.row_norms(n * C)
    st.counters[0] = n
    st.invalidate_shadow()


@pytest.mark.parametrize("C,bev", [(1, True), (2, False)])
@pytest.mark.parametrize("shards", [2, 3])
def test_tile_sharded_query_merge_equals_joint(pkg, C, bev, shards):
    """SURVEY 8(e), queries: row tiles of the grid on different ranks, local top-K, K-way merge (dist.merge_topk) == the query over
    one store holding every cell, bit for bit -- with duplicated rows (ties) and fewer than K candidates on some shards."""
    dmod = importlib.import_module("bsc-nav_b200.dist")
    gs, Dq, n = 96, 128, 3000
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=gs, floor_height=-1.0, map_height=1.0, token_dim=Dq, cache_size=C, bev=bev,
                          max_slots=4096)
    gen = torch.Generator("cuda").manual_seed(10 * C + shards)
    joint = pkg.SpatialStore(cfg)
    levels = joint.g.levels
    cells = torch.randperm(gs * gs * levels, device="cuda", generator=gen)[:n].to(torch.int32)
    base = torch.randn(200, Dq, device="cuda", generator=gen)
    feat = base[torch.randint(0, 200, (n * C,), device="cuda", generator=gen)]                  # many identical rows: ties
    tok_h = torch.randint(0, max(joint.g.maxh - joint.g.minh, 1), (n * C,), device="cuda", generator=gen).to(torch.int32)
    _install_cells(joint, cells, feat, tok_h)
    q = torch.cat([base[:3] + 0.05 * torch.randn(3, Dq, device="cuda", generator=gen), torch.randn(2, Dq, device="cuda", generator=gen)])
    row = (cells.long() // levels) // gs
    # uneven tiles: the last shard holds only the last 2 rows of the grid (fewer than K cells)
    owner = torch.where(row >= gs - 2, shards - 1, row % max(shards - 1, 1))
    for k in (100, 7):
        parts = []
        for r in range(shards):
            sel = (owner == r).nonzero().flatten()
            rows = (sel[:, None] * C + torch.arange(C, device="cuda")[None]).flatten()
            st = pkg.SpatialStore(cfg)
            _install_cells(st, cells[sel].contiguous(), feat[rows], tok_h[rows])
            parts.append(st.query(q, k=k) if r % 2 == 0 else st.query_batch(q, k=k))
        got = dmod.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                              torch.stack([p[2] for p in parts]), k)
        want = joint.query(q, k=k)
        for a, b, name in zip(got, want, ("sims", "positions", "counts")):
            assert torch.equal(a, b), f"k={k} {name} differ"


def test_two_rank_nccl_merge():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_rank_merge.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MERGE_OK" in out.stdout
