"""CPU, world_size 2, gloo: the host side of the multi-GPU path (shard plan, variable-length delta all-gather protocol,
rank-count independence of the winner rule)."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_delta(rank, n, dim=8):
    rng = np.random.default_rng(100 + rank)
    cell = rng.integers(0, 50, size=n).astype(np.int32)
    r2 = rng.uniform(0.1, 9.0, size=n).astype(np.float32)
    ordinal = (np.arange(n) + 1000 * rank).astype(np.uint64)
    key = (r2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | ordinal
    h = rng.integers(0, 200, size=n).astype(np.int32)
    feat = rng.standard_normal((n, dim)).astype(np.float32)
    return cell, key, h, feat


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("bsc-nav_b200.dist")
    n = 7 if rank == 0 else 12                      # ragged: different counts per rank
    cell, key, h, feat = _make_delta(rank, n)
    pad = 5                                          # rows beyond n must be ignored
    t = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((pad,) + a.shape[1:], a.dtype)]))
    g_cell, g_key, g_h, g_feat, counts = d.gather_deltas(t(cell), t(key.view(np.int64)), t(h), t(feat), n)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), cell=g_cell.numpy(), key=g_key.numpy(), h=g_h.numpy(), feat=g_feat.numpy(),
             counts=np.array(counts))
    # empty contribution on one rank
    e_cell, e_key, e_h, e_feat, e_counts = d.gather_deltas(t(cell), t(key.view(np.int64)), t(h), t(feat), 0 if rank == 1 else n)
    assert e_counts == [7, 0] and e_cell.shape[0] == 7
    dist.destroy_process_group()


def test_gather_deltas_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ("cell", "key", "h", "feat", "counts"):
        assert np.array_equal(r0[k], r1[k])                       # every rank ends with the same gathered deltas
    assert r0["counts"].tolist() == [7, 12]
    c0, k0, h0, f0 = _make_delta(0, 7)
    c1, k1, h1, f1 = _make_delta(1, 12)
    assert np.array_equal(r0["cell"], np.concatenate([c0, c1]))  # rank order, padding stripped
    assert np.array_equal(r0["key"].view(np.uint64), np.concatenate([k0, k1]))
    assert np.array_equal(r0["feat"], np.concatenate([f0, f1]))
    # winner rule is independent of how the rows were split across ranks
    d = importlib.import_module("bsc-nav_b200.dist")
    merged = d.merge_rule_host(r0["cell"], r0["key"].view(np.uint64), cache=3)
    swapped = d.merge_rule_host(np.concatenate([c1, c0]), np.concatenate([k1, k0]), cache=3)
    assert merged == swapped


def test_shard_plans():
    d = importlib.import_module("bsc-nav_b200.dist")
    for n, w in ((1000, 8), (10, 4), (3, 8), (0, 2)):
        shards = [d.shard_frames(n, r, w) for r in range(w)]
        assert sum(len(s) for s in shards) == n
        assert [i for s in shards for i in s] == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert d.shard_scenes(90, 3, 8) == list(range(3, 90, 8))
    assert sorted(i for r in range(4) for i in d.shard_scenes(10, r, 4)) == list(range(10))


class _FakeStore:
    """Stands in for SpatialStore.query in the CPU test: returns this rank's pre-computed local top-K."""

    def __init__(self, sims, pos, cnt):
        self.res = (sims, pos, cnt)

    def query(self, q, k=100, **_):
        return self.res


def _local_topk(rank, Q, K):
    g = torch.Generator().manual_seed(50 + rank)
    n = K if rank == 0 else K - 3                                # the second rank has fewer than K candidates
    sims = torch.full((Q, K), float("-inf"))
    pos = torch.full((Q, K, 3), -1, dtype=torch.int32)
    for qi in range(Q):
        s = torch.sort(torch.randn(n, generator=g), descending=True)[0]
        sims[qi, :n] = s
        pos[qi, :n] = torch.stack([torch.randperm(40, generator=g)[:n] * 2 + rank, torch.randint(0, 40, (n,), generator=g),
                                   torch.randint(0, 20, (n,), generator=g)], dim=1).to(torch.int32)   # rows of rank r: parity r
    return sims, pos, torch.full((Q,), n, dtype=torch.int32)


def _query_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("bsc-nav_b200.dist")
    s, p, c = d.query_sharded(_FakeStore(*_local_topk(rank, 3, 10)), torch.zeros(3, 8), k=10)
    np.savez(os.path.join(out_dir, f"q{rank}.npz"), s=s.numpy(), p=p.numpy(), c=c.numpy())
    dist.destroy_process_group()


def test_query_sharded_world2(tmp_path):
    port = _free_port()
    mp.spawn(_query_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "q0.npz"), np.load(tmp_path / "q1.npz")
    for k in ("s", "p", "c"):
        assert np.array_equal(r0[k], r1[k])                       # every rank ends with the same merged top-K
    (s0, p0, _), (s1, p1, _) = _local_topk(0, 3, 10), _local_topk(1, 3, 10)
    for qi in range(3):
        allv = [(float(s0[qi, j]), tuple(p0[qi, j].tolist())) for j in range(10)] + [(float(s1[qi, j]), tuple(p1[qi, j].tolist())) for j in range(7)]
        exp = sorted(allv, key=lambda t: (-t[0], t[1]))[:10]
        assert r0["c"][qi] == 10
        assert [tuple(v) for v in r0["p"][qi].tolist()] == [e[1] for e in exp]
        assert np.array_equal(r0["s"][qi], np.array([e[0] for e in exp], dtype=np.float32))


def test_merge_topk_against_brute_force():
    """K-way merge of per-rank results: per-cell max over ranks, similarity descending, ties by (row, col, h), padding past the
    count -- against a dictionary-and-sort statement, with duplicated cells across ranks, empty ranks and k < candidates."""
    d = importlib.import_module("bsc-nav_b200.dist")
    for seed in range(6):
        g = torch.Generator().manual_seed(seed)
        R, Q, K = [(3, 4, 10), (2, 1, 16), (4, 3, 5)][seed % 3]
        k = K if seed % 2 else max(1, K - 3)
        sims = torch.randn(R, Q, K, generator=g)
        if seed >= 3:
            sims = torch.round(sims * 2) / 2                        # coarse values: many exact ties
        pos = torch.randint(0, 5, (R, Q, K, 3), generator=g, dtype=torch.int32)   # small range: cells repeat across ranks
        cnt = torch.randint(0, K + 1, (R, Q), generator=g, dtype=torch.int32)
        cnt[:, 0] = 0 if seed == 0 else cnt[:, 0]                   # a query nobody has candidates for
        s, p, c = d.merge_topk(sims, pos, cnt, k)
        for qi in range(Q):
            best = {}
            for r in range(R):
                for j in range(int(cnt[r, qi])):
                    key, v = tuple(pos[r, qi, j].tolist()), float(sims[r, qi, j])
                    best[key] = max(best.get(key, float("-inf")), v)
            exp = sorted(best.items(), key=lambda kv: (-kv[1], kv[0]))[:k]
            assert int(c[qi]) == len(exp)
            assert [tuple(v) for v in p[qi, :len(exp)].tolist()] == [e[0] for e in exp]
            assert s[qi, :len(exp)].tolist() == [float(np.float32(e[1])) for e in exp]
            assert torch.all(s[qi, len(exp):] == float("-inf")) and torch.all(p[qi, len(exp):] == -1)


# ---- dist.fuse_scene_sharded / merge_scene_across_ranks: the collective protocol with a host stand-in for the device store
_ROW = np.dtype([("cell", "<i4"), ("h", "<i4"), ("key", "<u8"), ("feat", "<f4", (4,))])      # packed export row, 16 + 4 D bytes


def _frame_points(f, n_pts):
    """What frame f writes: (cell, key, h, token row) per point; keys {r^2 bits, ordinal} as the kernels build them."""
    rng = np.random.default_rng(7 + f % 3)                 # frames f and f + 3 see the same cells at the same r^2: ordinal ties
    cell = rng.integers(0, 40, size=n_pts).astype(np.int32)
    r2 = rng.uniform(0.2, 4.0, size=n_pts).astype(np.float32)
    ordinal = np.arange(n_pts, dtype=np.uint64) + np.uint64(f * n_pts)
    key = (r2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | ordinal
    feat = np.stack([np.full(4, f * 1000 + i, dtype=np.float32) for i in range(n_pts)])
    return cell, key, (cell % 7).astype(np.int32), feat


class _HostStore:
    """C = 1 store under the winner rule min (r^2, ordinal), with the packed export / merge interface of SpatialStore."""
    device = torch.device("cpu")

    def __init__(self):
        self.cells = {}
        self.next_ordinal = 0

    def put(self, cell, key, h, feat):
        for c, k, hh, ft in zip(cell.tolist(), key.tolist(), h.tolist(), feat):
            if c not in self.cells or k < self.cells[c][0]:
                self.cells[c] = (k, hh, ft.copy())

    def export_packed(self, ord_lo, ord_hi, max_out):
        rows = [(c, hh, k, ft) for c, (k, hh, ft) in sorted(self.cells.items()) if ord_lo <= (k & 0xFFFFFFFF) < ord_hi]
        out = np.zeros(max_out, dtype=_ROW)
        for i, r in enumerate(rows[:max_out]):
            out[i] = r
        n = torch.tensor([min(len(rows), max_out), len(rows)], dtype=torch.int32)
        return torch.from_numpy(out.view(np.uint8).reshape(max_out, _ROW.itemsize)), n

    def merge_packed(self, packed, n_segments, seg_capacity, seg_counts, count_stride=2, skip_segment=-1):
        rows = packed.numpy().reshape(-1).view(_ROW).reshape(n_segments, seg_capacity)
        for s in range(n_segments):
            if s != skip_segment:
                r = rows[s, :int(seg_counts[s * count_stride])]
                self.put(r["cell"], r["key"], r["h"], r["feat"])


class _HostMemory:
    """The calls fuse_scene_sharded makes on VoxelTokenMemory, recorded and applied to a _HostStore."""

    def __init__(self, h, w):
        import types
        self.args = types.SimpleNamespace(height=h, width=w)
        self._store = _HostStore()
        self.origin = None
        self.fed = []

    def begin_shard(self, origin_pose, first_frame):
        self.origin = np.asarray(origin_pose).copy()
        self._store.next_ordinal = first_frame * self.args.height * self.args.width

    def fuse_frames(self, rgb, depth, poses):
        n_pts = self.args.height * self.args.width
        for f in rgb.tolist():                                         # "rgb" carries the global frame numbers
            assert self._store.next_ordinal == f * n_pts               # points numbered by GLOBAL frame index on every rank
            self._store.put(*_frame_points(f, n_pts))
            self._store.next_ordinal += n_pts
            self.fed.append(f)
        assert len(poses) == len(rgb)

    def _join(self):
        pass


def _scene_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("bsc-nav_b200.dist")
    F = 11
    poses = np.arange(F * 7, dtype=np.float64).reshape(F, 7)
    mem = _HostMemory(2, 3)
    counts = d.fuse_scene_sharded(mem, poses, lambda f0, n: (np.arange(f0, f0 + n), None), batch=4, max_rows=64)
    assert np.array_equal(mem.origin, poses[0])                        # every shard hangs on the scene's first pose
    assert mem.fed == list(d.shard_frames(F, rank, world))
    cells = mem._store.cells
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), cell=np.array(sorted(cells)), key=np.array([cells[c][0] for c in sorted(cells)], dtype=np.uint64),
             feat=np.stack([cells[c][2] for c in sorted(cells)]), counts=np.array(counts))
    try:                                                                # a rank that produced more rows than max_rows is an error
        d.merge_scene_across_ranks(mem._store, 0, 2 ** 32, 3)
        raised = False
    except RuntimeError:
        raised = True
    assert raised
    dist.destroy_process_group()


def test_fuse_scene_sharded_world2(tmp_path):
    port = _free_port()
    mp.spawn(_scene_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    joint = _HostStore()
    for f in range(11):
        joint.put(*_frame_points(f, 6))
    want_cells = np.array(sorted(joint.cells))
    for r in range(2):
        got = np.load(tmp_path / f"s{r}.npz")
        assert np.array_equal(got["cell"], want_cells)
        assert np.array_equal(got["key"], np.array([joint.cells[c][0] for c in want_cells], dtype=np.uint64))
        assert np.array_equal(got["feat"], np.stack([joint.cells[c][2] for c in want_cells]))
        assert got["counts"].tolist() == np.load(tmp_path / "s0.npz")["counts"].tolist() and len(got["counts"]) == 2
