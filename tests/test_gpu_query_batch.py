"""Batched query path (csrc/query_batch.cu: tcgen05 candidate GEMM + exact fp32 re-scoring) against the exact single-query
path (k_query_slots): similarities, positions and counts must be IDENTICAL, bit for bit -- including ties (the same ViT
patch token is scattered to many cells, so duplicate rows are the norm), near-ties around the K-th value, filters, cells
with several tokens (C > 1) and empty rows, fewer than K candidates, more than 1024 queries (several GEMM passes), and the
overflow fall-back.  The exact path itself is checked against the oracle in test_gpu_memory.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _install(st, n, gen, dup_groups=0, empty_frac=0.0, scale=1.0):
    """Fill n slots of a store by hand: random rows (optionally groups of identical rows and empty rows)."""
    g = st.g
    C, D = g.cache, g.dim
    rows = n * C
    feat = torch.randn(rows, D, device="cuda", generator=gen) * scale
    if dup_groups:
        src = torch.randint(0, dup_groups, (rows,), device="cuda", generator=gen)
        feat = feat[:dup_groups][src]
    st.feat[:rows] = feat
    cells = torch.randperm(g.gs * g.gs * g.levels, device="cuda", generator=gen)[:n].to(torch.int32)
    st.occupied_ids.fill_(-1)
    st.slot_cell[:n] = cells
    st.occupied_ids[cells.long()] = torch.arange(n, dtype=torch.int32, device="cuda")
    key = torch.zeros(rows, dtype=torch.int64, device="cuda")
    if empty_frac:
        empty = torch.rand(rows, device="cuda", generator=gen) < empty_frac
        empty.view(n, C)[:, 0] = False                                  # every slot keeps at least one token
        key[empty] = -1                                                 # BSCNAV_KEY_EMPTY
    st.row_key[:rows] = key
    st.slot_keys[:rows] = key
    st.tok_h[:rows] = torch.randint(0, max(g.maxh - g.minh, 1), (rows,), device="cuda", generator=gen).to(torch.int32)
    st.row_norms(rows)
    st.counters[0] = n


def _same(a, b):
    for x, y, name in zip(a, b, ("sims", "positions", "counts")):
        assert torch.equal(x, y), f"{name} differ: {(x != y).sum().item()} entries"


@pytest.mark.parametrize("C,bev,n,D,dup", [(1, True, 5000, 256, 0), (1, True, 20000, 1024, 300), (3, False, 3000, 384, 0),
                                          (10, False, 700, 1024, 40), (2, True, 4097, 128, 0)])
@pytest.mark.parametrize("Q", [1, 7, 300, 1100])      # 1100: two GEMM passes (1024 queries per pass)
def test_batch_equals_exact(pkg, C, bev, n, D, dup, Q):
    cfg = pkg.StoreConfig(height=96, width=128, grid_size=160, token_dim=D, cache_size=C, bev=bev, floor_height=-1.0, map_height=1.0,
                          max_slots=25600)
    st = pkg.SpatialStore(cfg)
    gen = torch.Generator("cuda").manual_seed(C * 1000 + Q)
    _install(st, n, gen, dup_groups=dup, empty_frac=0.3 if C > 1 else 0.0)
    q = torch.randn(Q, D, device="cuda", generator=gen)
    q[::3] = st.feat[torch.randint(0, n * C, (len(q[::3]),), device="cuda", generator=gen)] + 0.1 * q[::3]   # some queries near stored rows
    for k in (100, 10):
        _same(st.query_batch(q, k=k), st.query(q, k=k))
    # region + floor filters
    kw = dict(region_radius=37.5, curr_grid=(80, 70, 9), floor_range=(3, 15))
    _same(st.query_batch(q, k=50, **kw), st.query(q, k=50, **kw))
    # a degenerate (empty) floor range from load_single_floor selects nothing -- it must not switch the filter off
    for path in (st.query, st.query_batch):
        _s, _p, cnt = path(q, k=50, floor_range=(9, 3))
        assert cnt.tolist() == [0] * Q


def test_batch_fewer_than_k_and_empty(pkg):
    cfg = pkg.StoreConfig(height=96, width=128, grid_size=64, token_dim=128, cache_size=1, bev=True)
    st = pkg.SpatialStore(cfg)
    gen = torch.Generator("cuda").manual_seed(5)
    q = torch.randn(4, 128, device="cuda", generator=gen)
    _same(st.query_batch(q, k=20), st.query(q, k=20))           # empty store
    _install(st, 13, gen)
    got = st.query_batch(q, k=20)
    _same(got, st.query(q, k=20))
    assert got[2].tolist() == [13] * 4 and (got[1][:, 13:] == -1).all()


def test_batch_near_ties_around_kth(pkg):
    """Thousands of cells within 1e-3 of each other around the K-th similarity: the margin of the candidate selection must
    keep every cell the exact arithmetic could rank in the top-K."""
    D, n = 256, 6000
    cfg = pkg.StoreConfig(height=96, width=128, grid_size=128, token_dim=D, cache_size=1, bev=True)
    st = pkg.SpatialStore(cfg)
    gen = torch.Generator("cuda").manual_seed(2)
    _install(st, n, gen)
    q = torch.randn(3, D, device="cuda", generator=gen)
    base = q[0] / q[0].norm()
    noise = torch.randn(n, D, device="cuda", generator=gen)
    noise -= (noise @ base)[:, None] * base                         # orthogonal to the query
    noise /= noise.norm(dim=1, keepdim=True)
    eps = torch.linspace(0.30, 0.302, n, device="cuda")             # cos = 1/sqrt(1+eps^2): spread of ~6e-4 over 6000 cells
    st.feat[:n] = base + eps[:, None] * noise
    st.row_norms(n)
    st.invalidate_shadow()
    _same(st.query_batch(q, k=100), st.query(q, k=100))


def test_batch_overflow_falls_back(pkg):
    """More identical rows than the batch path's candidate capacity: flagged with count -1 by the C ABI, re-issued through
    the exact path by the wrapper."""
    D, n = 128, 6000
    cfg = pkg.StoreConfig(height=96, width=128, grid_size=128, token_dim=D, cache_size=1, bev=True)
    st = pkg.SpatialStore(cfg)
    gen = torch.Generator("cuda").manual_seed(3)
    _install(st, n, gen, dup_groups=1)
    q = torch.randn(2, D, device="cuda", generator=gen)
    raw = st.query_batch(q, k=100, check=False)
    assert raw[2].tolist() == [-1, -1]
    _same(st.query_batch(q, k=100), st.query(q, k=100))


def test_batch_after_fuse_uses_fresh_shadow(pkg):
    """The shadow is rebuilt when the store changed."""
    from oracle import memory_port as mp
    H, W, gs, D = 96, 128, 64, 128
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=gs, token_dim=D, cache_size=2, bev=False, floor_height=-2.0, map_height=2.0)
    st = pkg.SpatialStore(cfg)
    chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    gen = torch.Generator("cuda").manual_seed(8)
    q = torch.randn(9, D, device="cuda", generator=gen)
    for f in range(3):
        _, depth, pose = mp.synthetic_frame(f, H, W, n_frames=8)
        tokens = torch.randn(1, 256, D, device="cuda", generator=gen)
        st.fuse(torch.from_numpy(depth[None]).cuda(), chain(pose)[None], tokens)
        _same(st.query_batch(q, k=30), st.query(q, k=30))


def test_batch_full_size_planted(pkg):
    """Config-3/4 shape: 262 144 stored tokens, D = 1024, 256 queries in one pass, K = 100."""
    gs, D, K, Q = 512, 1024, 100, 256
    cfg = pkg.StoreConfig(height=480, width=640, grid_size=gs, token_dim=D, cache_size=1, bev=True)
    st = pkg.SpatialStore(cfg)
    n = gs * gs
    gen = torch.Generator("cuda").manual_seed(7)
    st.feat.copy_(torch.randn(n, D, device="cuda", generator=gen))
    q = torch.randn(Q, D, device="cuda", generator=torch.Generator("cuda").manual_seed(11))
    rng = np.random.default_rng(3)
    for i in range(0, Q, 16):
        idx = torch.from_numpy(rng.choice(n, size=K, replace=False)).cuda()
        st.feat[idx] = q[i] + 0.05 * torch.randn(K, D, device="cuda", generator=gen)
    st.slot_cell.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
    st.occupied_ids.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
    st.row_key.zero_(); st.slot_keys.zero_(); st.tok_h.fill_(100)
    st.row_norms(n)
    st.counters[0] = n
    got = st.query_batch(q, k=K)
    assert (got[2] == K).all()
    for i0 in range(0, Q, 8):
        ref = st.query(q[i0:i0 + 8], k=K)
        _same(tuple(t[i0:i0 + 8] for t in got), ref)
