"""Memory-directory interchange: the reference's HDF5 token layout (feat.h5df: group grid_{row}_{col}_{h} with `features` and
`distances`, memory_2.py:335-354) written from and read back into the device store, through an in-memory stand-in for h5py
(the module is not part of this image; io.export_h5 / io.import_h5 take it as an argument)."""
import importlib
import types

import numpy as np
import pytest
import torch

from oracle import memory_port as mp

pytestmark = pytest.mark.gpu

_FILES = {}


class _DS:
    def __init__(self, data):
        self._a = np.array(data, copy=True)
    shape = property(lambda self: self._a.shape)

    def __getitem__(self, k):
        return self._a[k]


class _Group(dict):
    def create_dataset(self, name, data=None, maxshape=None, chunks=None):
        self[name] = _DS(data)
        return self[name]


class _File:
    def __init__(self, path, mode="r"):
        self._g = _FILES.setdefault(path, {})

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, k):
        return self._g[k]

    def keys(self):
        return sorted(self._g)            # HDF5 iterates names in sorted order

    def create_group(self, name):
        assert name not in self._g
        self._g[name] = _Group()
        return self._g[name]


fake_h5 = types.SimpleNamespace(File=_File)


def test_h5_roundtrip(pkg, tmp_path):
    io = importlib.import_module("bsc-nav_b200.io")
    H, W, gs, D = 96, 128, 64, 128
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=gs, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=3, bev=False)
    chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    frames = [mp.synthetic_frame(f, H, W, n_frames=10) for f in range(3)]
    depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
    tok = torch.randn(3, 256, D, device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    a = pkg.SpatialStore(cfg)
    a.fuse(depth, np.stack([chain(f[2]) for f in frames]), tok)
    path = str(tmp_path)
    io.export_h5(types.SimpleNamespace(store=a), path, fake_h5)
    np.save(path + "/occupied_ids.npy", a.occupied_ids.cpu().numpy().reshape(gs, gs, a.g.levels))
    # the file holds the reference's layout: one group per voxel, best-first tokens, float32 distances
    files = _FILES[path + "/feat.h5df"]
    sa = a.snapshot()
    assert len(files) == a.count() and set(files) == {f"grid_{k[0]}_{k[1]}_{k[2]}" for k in sa}
    k0 = next(iter(sa))
    g0 = files[f"grid_{k0[0]}_{k0[1]}_{k0[2]}"]
    assert g0["features"].shape == (len(sa[k0]), D) and g0["distances"][:].dtype == np.float32
    b = pkg.SpatialStore(cfg)
    io.import_h5(types.SimpleNamespace(store=b), path, fake_h5)
    sb = b.snapshot()
    assert b.count() == a.count() and torch.equal(a.slot_cell[: a.count()], b.slot_cell[: b.count()])
    assert sa.keys() == sb.keys()
    for k in sa:
        assert [t[0] for t in sa[k]] == [t[0] for t in sb[k]] and [t[2] for t in sa[k]] == [t[2] for t in sb[k]]
        assert all(np.array_equal(x[3], y[3]) for x, y in zip(sa[k], sb[k]))
    q = torch.randn(5, D, device="cuda", generator=torch.Generator("cuda").manual_seed(6))
    ra, rb = a.query(q, k=40), b.query(q, k=40)
    assert all(torch.equal(x, y) for x, y in zip(ra, rb))
