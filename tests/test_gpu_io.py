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


def _small_memory(pkg, colour=False, path=None, load=False):
    synth = importlib.import_module("bsc-nav_b200.synth")
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=1, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=1), max_batch=3)
    args = synth.Args(height=96, width=128, grid_size=64, depth_sample_rate=1, floor_height=-2.0, map_height=2.0, load_memory_path=path)
    mem = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=3, max_slots=65536, colour=colour)
    if load:
        mem.load_memory()
    return mem, synth


def test_real_hdf5_file_interchange(pkg, tmp_path):
    """A real feat.h5df on disk (written by h5lite, the package's implementation of the file family h5py produces): save, save
    AGAIN into the same directory (must not fail on existing groups), then load the directory the way a memory written
    by the reference is loaded -- tokens from feat.h5df, slot numbering from occupied_ids.npy -- and query both."""
    io = importlib.import_module("bsc-nav_b200.io")
    h5 = importlib.import_module("bsc-nav_b200.h5lite")
    mem, synth = _small_memory(pkg)
    frames = [synth.synthetic_frame(f, 96, 128, n_frames=10) for f in range(3)]
    mem.fuse_frames(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), np.stack([f[2] for f in frames]))
    path = str(tmp_path / "scene")
    mem.save_memory(path)
    mem.save_memory(path)
    snap = mem.store.snapshot()
    with h5.File(path + "/feat.h5df", "r") as f:
        assert len(f.keys()) == mem.max_id
        k0 = next(iter(snap))
        grp = f[f"grid_{k0[0]}_{k0[1]}_{k0[2]}"]
        assert grp["features"].shape == (len(snap[k0]), 128) and grp["distances"][:].dtype == np.float32
        assert np.array_equal(grp["features"][:], np.stack([t[3] for t in snap[k0]]))
    for name in [n for n in __import__("os").listdir(path) if n.startswith("bscnav_")]:
        __import__("os").remove(path + "/" + name)                      # what is left is the reference's own file set
    mem2, _ = _small_memory(pkg, path=path, load=True)
    assert mem2.max_id == mem.max_id and np.array_equal(mem2.grid_rgb_pos, mem.grid_rgb_pos)
    q = torch.randn(4, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(6))
    for a, b in zip(mem.store.query(q, k=30), mem2.store.query(q, k=30)):
        assert torch.equal(a, b)


def test_reference_directory_with_more_slots_than_groups(pkg, tmp_path):
    """The reference numbers a voxel even when its token is dropped at a buffer flush (memory_2.py:880-894): max_id and
    occupied_ids can name slots that have no group in feat.h5df.  Such slots load as empty voxels."""
    h5 = importlib.import_module("bsc-nav_b200.h5lite")
    io = importlib.import_module("bsc-nav_b200.io")
    gs, levels, D = 64, 40, 128
    path = str(tmp_path)
    pos = np.array([[3, 4, 5], [10, 11, 12], [20, 21, 22], [30, 31, 32]], dtype=np.int32)      # slot 2 has no tokens on disk
    occ = np.full((gs, gs, levels), -1, dtype=np.int32)
    for s, (r, c, h) in enumerate(pos):
        occ[r, c, h] = s
    np.save(path + "/occupied_ids.npy", occ); np.save(path + "/grid_rgb_pos.npy", pos); np.save(path + "/max_id.npy", np.array(4))
    np.save(path + "/map_height.npy", np.array([-20, 20]))
    rng = np.random.default_rng(1)
    feats = {s: rng.standard_normal((m, D)).astype(np.float32) for s, m in ((0, 2), (1, 5), (3, 1))}
    with h5.File(path + "/feat.h5df", "w") as f:
        for s, a in feats.items():
            g = f.create_group(f"grid_{pos[s][0]}_{pos[s][1]}_{pos[s][2]}")
            g.create_dataset("features", data=a, maxshape=(None, D), chunks=True)
            g.create_dataset("distances", data=np.arange(len(a), dtype=np.float32) + s, maxshape=(None,), chunks=True)
    mem, _ = _small_memory(pkg, path=path, load=True)
    assert mem.max_id == 4 and np.array_equal(mem.grid_rgb_pos[:, :2], pos[:, :2])
    snap = mem.store.snapshot()
    assert snap[(20, 21, 22)] == [] and len(snap[(10, 11, 12)]) == 3 and len(snap[(3, 4, 5)]) == 2     # C = 3 best of 5
    assert np.array_equal(snap[(10, 11, 12)][0][3], feats[1][0])
    q = torch.from_numpy(feats[3][0:1]).cuda()
    _, p, c = mem.store.query(q, k=10)
    assert int(c[0]) == 3 and p[0, 0].tolist() == [30, 31, 32]                                   # the empty voxel takes no part


def test_colour_channels_survive_save_and_load(pkg, tmp_path):
    """load_memory restores grid_rgb / weight / cv_map (memory_2.py:191-193) and continued fusing carries on from
    them: build 4 frames at once == build 2, save, load, build 2 more."""
    mem_a, synth = _small_memory(pkg, colour=True)
    frames = [synth.synthetic_frame(f, 96, 128, n_frames=10) for f in range(4)]
    rgb, depth, poses = (np.stack([f[i] for f in frames]) for i in range(3))
    for i in range(4):
        mem_a.fuse_frames(rgb[i:i + 1], depth[i:i + 1], poses[i:i + 1])
    mem_b, _ = _small_memory(pkg, colour=True)
    for i in range(2):
        mem_b.fuse_frames(rgb[i:i + 1], depth[i:i + 1], poses[i:i + 1])
    path = str(tmp_path / "scene")
    mem_b.save_memory(path)
    mem_c, _ = _small_memory(pkg, colour=True, path=path, load=True)
    assert np.array_equal(mem_c.grid_rgb, mem_b.grid_rgb) and np.array_equal(mem_c.weight, mem_b.weight)
    assert np.array_equal(mem_c.cv_map, mem_b.cv_map) and np.array_equal(mem_c.max_height, mem_b.max_height)
    assert not mem_c.is_unknown(*mem_c.grid_rgb_pos[0][:2])
    mem_c._chain.inv_init_base_tf = mem_b._chain.inv_init_base_tf          # the scene's origin travels with the caller (memory_2.py:844-846)
    for i in range(2, 4):
        mem_c.fuse_frames(rgb[i:i + 1], depth[i:i + 1], poses[i:i + 1])
    assert mem_c.max_id == mem_a.max_id
    assert np.array_equal(mem_c.grid_rgb, mem_a.grid_rgb) and np.array_equal(mem_c.weight, mem_a.weight)
    assert np.array_equal(mem_c.cv_map, mem_a.cv_map) and np.array_equal(mem_c.max_height, mem_a.max_height)
    with pytest.raises(ValueError, match="height range"):
        other = importlib.import_module("bsc-nav_b200.synth").Args(height=96, width=128, grid_size=64, depth_sample_rate=1,
                                                                   floor_height=-3.0, map_height=1.0, load_memory_path=path)     # same 40 levels, shifted
        m = pkg.VoxelTokenMemory(other, preload_dino=mem_a.dinov2, need_diffusion=False, cache_size=3, max_slots=65536)
        m.load_memory()
