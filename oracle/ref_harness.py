"""TEST INFRASTRUCTURE ONLY -- executes the *unmodified* reference source as the ground-truth oracle.

Only usable where /root/reference is mounted (the build container).  It is used by
`oracle/make_golden.py` to generate the fixtures under tests/golden/ and by the
`-m "not gpu"` tests to validate the numpy restatement in `oracle/memory_port.py`.
Nothing in the product package, the `-m gpu` tests, smoke() or bench.py imports it.

How: `import memory_2` fails here (h5py, habitat_sim, diffusers, ultralytics, open3d,
matplotlib are absent), so the hot-path functions are selected by name from the parsed
AST of /root/reference/{utils.py,memory_2.py} and exec'ed, source untouched, in a
namespace that provides numpy/torch/scipy plus a small in-memory stand-in for h5py.File
(SURVEY.md section 8c).  Only instance attributes and injected modules differ from a
real run; no reference line is edited or copied into this repo.
"""
from __future__ import annotations

import ast
import gc
import json
import math
import os
import random
import re
import time
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BSCNAV_REFERENCE_ROOT", "/root/reference")

UTILS_FUNCS = (
    "cvt_pose_vec2tf", "get_sim_cam_mat", "get_sim_cam_mat_with_fov", "depth2pc",
    "transform_pc", "base_pos2grid_id_3d", "project_point",
)
MEMORY_METHODS = (
    "_init_cache", "_reinit_cache", "_get_patch_token", "_backproject_depth", "_out_of_range",
    "obs2voxeltoken", "update_memory_dist_base", "get_total_token_count", "voxel_localized",
    "long_memory_filter",
)
AGENT_METHODS = ("weighted_cluster_centers",)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "memory_2.py"))


# --------------------------------------------------------------------------------------
# in-memory stand-in for the subset of h5py the reference touches
# (memory_2.py:312-358, 623-663): File(path, mode) context manager, `name in f`,
# iteration / keys() in sorted (HDF5 B-tree) order, create_group, group[...],
# create_dataset(data=, maxshape=, chunks=), dataset.shape / resize / [] get+set.
# --------------------------------------------------------------------------------------
class _Dataset:
    def __init__(self, data):
        self._a = np.array(data, copy=True)

    @property
    def shape(self):
        return self._a.shape

    def resize(self, shape):
        new = np.zeros(shape, dtype=self._a.dtype)
        n = min(shape[0], self._a.shape[0])
        new[:n] = self._a[:n]
        self._a = new

    def __getitem__(self, k):
        return self._a[k]

    def __setitem__(self, k, v):
        self._a[k] = v


class _Group(dict):
    def create_dataset(self, name, data=None, maxshape=None, chunks=None):
        ds = _Dataset(data)
        self[name] = ds
        return ds


_FILES: dict = {}


class _File:
    def __init__(self, path, mode="r"):
        self._g = _FILES.setdefault(path, {})

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __contains__(self, k):
        return k in self._g

    def __getitem__(self, k):
        return self._g[k]

    def __iter__(self):
        return iter(sorted(self._g))

    def keys(self):
        return sorted(self._g)

    def create_group(self, name):
        g = _Group()
        self._g[name] = g
        return g


fake_h5py = types.SimpleNamespace(File=_File)


def fake_h5_files() -> dict:
    """path -> {group_name: {'features': _Dataset, 'distances': _Dataset}}"""
    return _FILES


def reset_fake_h5():
    _FILES.clear()


# --------------------------------------------------------------------------------------
# AST extraction
# --------------------------------------------------------------------------------------
def _parse(name):
    with open(os.path.join(REFERENCE_ROOT, name), "r", encoding="utf-8") as f:
        return ast.parse(f.read(), filename=name)


def load_utils_namespace() -> dict:
    """Reference geometry helpers (utils.py:133-214), executed as written."""
    from scipy.spatial.transform import Rotation as R

    tree = _parse("utils.py")
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in UTILS_FUNCS]
    assert {n.name for n in nodes} == set(UTILS_FUNCS)
    ns = {"np": np, "R": R, "math": math}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), "reference/utils.py", "exec"), ns)
    return ns


def load_memory_class(identity_shuffle: bool = False):
    """Reference VoxelTokenMemory restricted to the hot-path methods (memory_2.py:38), executed as written.

    identity_shuffle=True swaps `np.random.shuffle` for a no-op *inside the exec namespace only*
    (memory_2.py:747-748 permutes even at rate=1); otherwise seed with np.random.seed first.
    """
    import torch
    import torch.nn.functional as F
    from torchvision import transforms

    ns = load_utils_namespace()
    np_mod = np
    if identity_shuffle:
        # shallow proxy of the numpy module whose random.shuffle leaves the array in place
        np_mod = types.ModuleType("numpy_identity_shuffle")
        np_mod.__dict__.update(np.__dict__)
        rnd = types.ModuleType("numpy_random_identity")
        rnd.__dict__.update(np.random.__dict__)
        rnd.shuffle = lambda a: None
        np_mod.random = rnd
    ns.update({
        "np": np_mod, "torch": torch, "F": F, "math": math, "random": random, "time": time, "re": re,
        "transforms": transforms, "h5py": fake_h5py, "gc": gc, "os": os, "json": json,
    })
    tree = _parse("memory_2.py")
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VoxelTokenMemory")
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in MEMORY_METHODS]
    assert {n.name for n in keep} == set(MEMORY_METHODS)
    new_cls = ast.ClassDef(name="VoxelTokenMemory", bases=[], keywords=[], body=keep, decorator_list=[])
    if hasattr(new_cls, "type_params"):
        new_cls.type_params = []
    mod = ast.Module(body=[new_cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, "reference/memory_2.py", "exec"), ns)
    return ns["VoxelTokenMemory"], ns


def load_agent_helpers():
    """weighted_cluster_centers (objnav_benchmark.py:477-495), executed as written."""
    from sklearn.cluster import DBSCAN

    tree = _parse("objnav_benchmark.py")
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GESObjectNavRobot")
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in AGENT_METHODS]
    new_cls = ast.ClassDef(name="GESObjectNavRobot", bases=[], keywords=[], body=keep, decorator_list=[])
    if hasattr(new_cls, "type_params"):
        new_cls.type_params = []
    mod = ast.Module(body=[new_cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "DBSCAN": DBSCAN}
    exec(compile(mod, "reference/objnav_benchmark.py", "exec"), ns)
    return ns["GESObjectNavRobot"]


def default_args(**over):
    """argparse defaults the memory ctor reads (args.py:24-67), overridable."""
    a = dict(
        height=680, width=680, query_height=224, query_width=224, cell_size=0.1, grid_size=1000,
        depth_sample_rate=1000, min_depth=0.1, max_depth=10, floor_height=-10.0, map_height=10.0,
        sensor_height=1.5, base_forward_axis=[0, 0, -1], base_left_axis=[-1, 0, 0], base_up_axis=[0, 1, 0],
        base2cam_rot=[1, 0, 0, 0, -1, 0, 0, 0, -1], load_single_floor=False, imagenary_num=3,
    )
    a.update(over)
    return types.SimpleNamespace(**a)


def make_reference_memory(args, vit, *, patch=14, token_dim=1024, iter_size=50000, cache_size=10,
                          identity_shuffle=False, feat_path="mem://oracle/feat.h5df"):
    """Instance built with __new__ + the attributes the ctor would set (memory_2.py:78-137), device='cpu'.

    `vit` is any object exposing forward_features(x) -> {'x_norm_patchtokens': (B, n_patch, D)}:
    the same plug-point (`preload_dino`, memory_2.py:43-46) the product engine uses.
    """
    from torchvision import transforms

    cls, ns = load_memory_class(identity_shuffle=identity_shuffle)
    m = cls.__new__(cls)
    m.args = args
    m.device = "cpu"
    m.dinov2 = vit
    m.transform = transforms.Compose([
        transforms.Resize((args.query_height, args.query_width)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    m.transform_ = transforms.Compose([
        transforms.Resize((args.query_height, args.query_width)),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    m.inv_init_base_tf = []
    m.patch_h = m.patch_w = patch
    m.n_patch_w = args.query_width // patch
    m.n_patch_h = args.query_height // patch
    m.base_transform = np.eye(4)
    m.base_transform[0, :3] = args.base_forward_axis
    m.base_transform[1, :3] = args.base_left_axis
    m.base_transform[2, :3] = args.base_up_axis
    m.base2cam_tf = np.eye(4)
    m.base2cam_tf[:3, :3] = np.array([args.base2cam_rot]).reshape((3, 3))
    m.base2cam_tf[1, 3] = args.sensor_height
    m.cs = args.cell_size
    m.gs = args.grid_size
    m.depth_sample_rate = args.depth_sample_rate
    m.cv_map = np.zeros((m.gs, m.gs, 3), dtype=np.uint8)
    m.max_height = np.full((m.gs, m.gs), -np.inf)
    m.calib_mat = ns["get_sim_cam_mat_with_fov"](args.height, args.width, fov=90)
    m.min_depth = args.min_depth
    m.max_depth = args.max_depth
    m.token_dim = token_dim
    m.iter_size = iter_size
    m.cache_size = cache_size
    m.floor_height = args.floor_height
    m.map_height = args.map_height
    m.maxh = int(m.map_height / m.cs)
    m.minh = int(m.floor_height / m.cs)
    (m.grid_feat, m.grid_feat_pos, m.grid_rgb_pos, m.grid_feat_dis, m.weight, m.occupied_ids,
     m.grid_rgb, m.max_id, m.iter_id, m.base_height) = m._init_cache()
    m.long_memory_dict = []
    m.feat_path = feat_path
    m._ns = ns
    return m


def flush_trimmed(m):
    """update_memory_dist_base() on exactly the rows written so far (harness rule):
    the reference iterates the whole 50 000-row buffer incl. unused zero rows (memory_2.py:331)."""
    n = m.iter_id
    m.grid_feat = m.grid_feat[:n]
    m.grid_feat_pos = m.grid_feat_pos[:n]
    m.grid_feat_dis = m.grid_feat_dis[:n]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m.update_memory_dist_base()


def stored_tokens(feat_path="mem://oracle/feat.h5df"):
    """{(row, col, h): (features (M, D) f32, distances (M,) f32)} from the stand-in HDF5 store."""
    out = {}
    for name, g in _FILES.get(feat_path, {}).items():
        _, r, c, h = name.split("_")
        out[(int(r), int(c), int(h))] = (np.array(g["features"][:]), np.array(g["distances"][:]))
    return out
