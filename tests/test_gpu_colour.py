"""Colour / occupancy side channels (csrc/colour.cu) against the reference's own arrays (tests/golden/*.npz were written by
the unmodified obs2voxeltoken, memory_2.py:888-903): grid_rgb, weight, cv_map and max_height must be IDENTICAL -- the uint8
cast after every update makes the colour mean a sequential chain, reproduced via a stable sort by voxel."""
import numpy as np
import pytest
import torch

from oracle import make_golden as mg
from oracle import memory_port as mp

pytestmark = pytest.mark.gpu


def _run(pkg, name, g, batched):
    ci = mg.case_inputs(name)
    a = ci["args"]
    tok = g["patch_tokens"]
    Fn, D = tok.shape[0], tok.shape[3]
    Dp = (D + 127) // 128 * 128
    tokp = np.zeros((Fn, tok.shape[1] * tok.shape[2], Dp), np.float32)
    tokp[:, :, :D] = tok.reshape(Fn, -1, D)
    cfg = pkg.StoreConfig(height=a["height"], width=a["width"], grid_size=a["grid_size"], floor_height=a.get("floor_height", -10.0),
                          map_height=a.get("map_height", 10.0), n_patch_h=tok.shape[2], n_patch_w=tok.shape[1], token_dim=Dp,
                          cache_size=10, bev=False, colour=True)
    st = pkg.SpatialStore(cfg)
    chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    depth = torch.from_numpy(np.stack([f[1] for f in ci["frames"]])).cuda()
    rgb = torch.from_numpy(np.stack([np.ascontiguousarray(f[0]) for f in ci["frames"]])).cuda()
    poses = np.stack([chain(f[2]) for f in ci["frames"]])
    tokens = torch.from_numpy(tokp).cuda()
    sidx = torch.from_numpy(g["sample_idx"].astype(np.int32)).cuda() if "sample_idx" in g else None
    if batched:
        st.fuse(depth, poses, tokens, sample_idx=sidx, rgb=rgb)
    else:
        for f in range(Fn):
            st.fuse(depth[f:f + 1].contiguous(), poses[f:f + 1], tokens[f:f + 1].contiguous(),
                    sample_idx=None if sidx is None else sidx[f:f + 1].contiguous(), rgb=rgb[f:f + 1].contiguous())
    return st


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("name", ["cfg1", "two_pose", "sampled"])
def test_colour_channels_match_reference(pkg, golden, name, batched):
    g = golden(name)
    st = _run(pkg, name, g, batched)
    grid_rgb, weight, cv_map, max_height = st.colour_views()
    assert st.count() == int(g["max_id"])
    assert np.array_equal(st.positions(), g["grid_rgb_pos"])
    assert np.array_equal(max_height, g["max_height"])
    assert np.array_equal(cv_map, g["cv_map"])
    assert np.array_equal(weight.view(np.uint32), g["weight"].view(np.uint32)), f"{(weight != g['weight']).sum()} weights differ"
    assert np.array_equal(grid_rgb, g["grid_rgb"]), f"{(grid_rgb != g['grid_rgb']).any(axis=1).sum()} colours differ"


def test_colour_dense_long_chains_match_port(pkg):
    """Dense 128 x 160 frames from nearby poses: hundreds of points per voxel, chains that cross batch boundaries."""
    H, W, gs = 128, 160, 64
    p = mp.MemoryParams(height=H, width=W, grid_size=gs, floor_height=-2.0, map_height=2.0, token_dim=128, cache_size=2, bev=False)
    om = mp.OracleMemory(p)
    cfg = pkg.StoreConfig(height=H, width=W, grid_size=gs, floor_height=-2.0, map_height=2.0, token_dim=128, cache_size=2, bev=False,
                          colour=True)
    st = pkg.SpatialStore(cfg)
    chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    frames = [mp.synthetic_frame(f, H, W, n_frames=40) for f in range(5)]
    Ts = [chain(f[2]) for f in frames]
    tok = np.zeros((16, 16, 128), np.float32)
    for f, (rgb, depth, _pose) in enumerate(frames):
        om.fuse_frame(rgb, depth, Ts[f], tok, frame_id=f, colour=True)
    tokens = torch.zeros(1, 256, 128, device="cuda")
    for lo, hi in ((0, 2), (2, 5)):                     # two batches
        depth = torch.from_numpy(np.stack([f[1] for f in frames[lo:hi]])).cuda()
        rgb = torch.from_numpy(np.stack([f[0] for f in frames[lo:hi]])).cuda()
        st.fuse(depth, np.stack(Ts[lo:hi]), tokens.expand(hi - lo, -1, -1).contiguous(), rgb=rgb)
    grid_rgb, weight, cv_map, max_height = st.colour_views()
    assert st.count() == len(om.grid_rgb)
    assert np.array_equal(max_height, om.max_height) and np.array_equal(cv_map, om.cv_map)
    ref_w = np.array(om.weight, dtype=np.float32)
    ref_rgb = np.stack(om.grid_rgb)
    # exp() on the device and in numpy may differ in the last bit of alpha: allow the rare 1-ulp / 1-level consequence
    assert (weight.view(np.int32) - ref_w.view(np.int32)).__abs__().max() <= 1
    diff = np.abs(grid_rgb.astype(np.int32) - ref_rgb.astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
