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
