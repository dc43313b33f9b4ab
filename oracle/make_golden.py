"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_harness.py).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
The fixtures pin oracle/memory_port.py (CPU tests) and the CUDA path (GPU tests) to the reference's own outputs.
Inputs are regenerated from seeds by the tests through the same helpers used here; everything order- or
RNG-dependent in the reference (np.random.shuffle sampling, memory_2.py:747-748) is stored explicitly.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import memory_port as mp  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle import vit_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class TableViT:
    """Deterministic stand-in feature extractor: a fixed (n_patch, D) token table + the image mean (so that
    the pre-processing path influences the result).  Same plug-point as `preload_dino`."""

    def __init__(self, n_patch, dim, seed):
        g = torch.Generator().manual_seed(seed)
        self.table = torch.randn(1, n_patch, dim, generator=g)

    def forward_features(self, x):
        return {"x_norm_patchtokens": self.table + x.mean(dim=(1, 2, 3)).view(-1, 1, 1)}


def case_inputs(name):
    """Shared by the generator and the tests: returns dict(args kwargs, frames [(rgb, depth, pose)], extras)."""
    if name == "cfg1":       # BASELINE.json configs[0]: single 128x128 frame, 64x64 grid, ViT-B/32 (7x7 patches)
        rgb = np.random.default_rng(0).integers(0, 256, size=(128, 128, 4), dtype=np.uint8)
        v, u = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")   # smooth surface + noise: stays under the
        depth = (1.5 + 0.5 * np.sin(u / 20.0) * np.cos(v / 15.0)             # reference's gs*gs slot capacity (memory_2.py:715-718)
                 + np.random.default_rng(1).uniform(-0.02, 0.02, size=(128, 128))).astype(np.float32)
        pose = np.array([0, 0, 0, 0, 0, 0, 1.0])
        return dict(args=dict(height=128, width=128, grid_size=64, depth_sample_rate=1, floor_height=-2.0,
                              map_height=2.0), frames=[(rgb, depth, pose)], patch=32, vit="clip-vit-b32")
    if name == "two_pose":   # non-square frames, two poses, 16x16 token table
        frames = [mp.synthetic_frame(f, 96, 128, n_frames=6) for f in (0, 1)]
        return dict(args=dict(height=96, width=128, grid_size=64, depth_sample_rate=1, floor_height=-2.0,
                              map_height=2.0), frames=frames, patch=14, vit=("table", 256, 32, 5))
    if name == "sampled":    # reference defaults (680^2, gs=1000, rate=1000, heights [-10,10)), seeded permutation
        frames = [mp.synthetic_frame(f, 680, 680, n_frames=12) for f in (0, 1, 2)]
        return dict(args=dict(height=680, width=680, grid_size=1000, depth_sample_rate=1000), frames=frames,
                    patch=14, vit=("table", 256, 32, 9), shuffle_seed=7)
    if name == "tilted":     # rolled / tilted poses (all nine rotation entries non-zero), sensor holes, points at / beyond both
        rng = np.random.default_rng(21)                                    # depth bounds: the erasure cases of the path
        frames = []
        for f in range(3):
            rgb, depth, pose = mp.synthetic_frame(f, 96, 128, n_frames=6)
            depth = depth.copy()
            depth[rng.random(depth.shape) < 0.08] = 0.0                    # holes
            depth[rng.random(depth.shape) < 0.04] = 25.0                   # beyond max_depth
            depth[rng.random(depth.shape) < 0.02] = np.float32(0.1)        # float32(0.1) > 0.1: just inside the strict bound
            depth[rng.random(depth.shape) < 0.02] = np.float32(10.0)       # exactly max_depth: dropped
            # a general rotation near the circle pose, built with +, *, / and sqrt only (correctly rounded everywhere), so the
            # GPU box regenerates bit-identical inputs whatever its libm / SIMD dispatch does with sin and cos
            q = pose[3:] + rng.uniform(-0.3, 0.3, size=4)
            q = q / math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
            pose = np.concatenate([pose[:3] + rng.uniform(-0.2, 0.2, size=3), q])
            frames.append((rgb, depth, pose))
        return dict(args=dict(height=96, width=128, grid_size=64, depth_sample_rate=1, floor_height=-2.0,
                              map_height=2.0), frames=frames, patch=14, vit=("table", 256, 32, 13))
    raise KeyError(name)


def build_vit(spec):
    if isinstance(spec, tuple):
        _, n_patch, dim, seed = spec
        return TableViT(n_patch, dim, seed), dim
    vc = importlib.import_module("bsc-nav_b200.vit_config")
    cfg = vc.CONFIGS[spec]
    return vit_ref.TorchViT(vc.random_weights(cfg, seed=0), cfg), cfg.dim


def run_reference(name):
    """name: a fixture name of case_inputs, or a case dict of the same shape (live cross-checks in the CPU tests)."""
    ci = case_inputs(name) if isinstance(name, str) else name
    vit, dim = build_vit(ci["vit"])
    args = rh.default_args(**ci["args"])
    rh.reset_fake_h5()
    seeded = "shuffle_seed" in ci
    m = rh.make_reference_memory(args, vit, patch=ci["patch"], token_dim=dim, iter_size=400000, cache_size=10 ** 9,
                                 identity_shuffle=not seeded)
    out = {}
    perms = []
    tokens = []
    for f, (rgb, depth, pose) in enumerate(ci["frames"]):
        if seeded:
            np.random.seed(ci["shuffle_seed"] + f)
            perm = np.arange(depth.size)
            np.random.shuffle(perm)
            perms.append(perm[:: args.depth_sample_rate].copy())
            np.random.seed(ci["shuffle_seed"] + f)      # same stream for the reference's own shuffle
        tokens.append(m._get_patch_token(np.array(rgb[:, :, :3])).numpy())
        m.obs2voxeltoken({"rgb": rgb, "depth": depth}, pose)
    n = m.iter_id
    out["patch_tokens"] = np.stack(tokens).astype(np.float32)            # (F, n_patch_w, n_patch_h, D)
    out["feat_pos"] = m.grid_feat_pos[:n].copy()
    out["feat_dis"] = m.grid_feat_dis[:n].copy()
    feat = m.grid_feat[:n].copy()
    out["max_id"] = np.array(m.max_id)
    out["grid_rgb_pos"] = m.grid_rgb_pos[: m.max_id].copy()
    out["grid_rgb"] = m.grid_rgb[: m.max_id].copy()
    out["weight"] = m.weight[: m.max_id].copy()
    out["cv_map"] = m.cv_map.copy()
    out["max_height"] = m.max_height.copy()
    out["minh_maxh"] = np.array([m.minh, m.maxh])
    if perms:
        out["sample_idx"] = np.stack(perms)
    # which token each buffer row holds (frame, ppy, ppx): recovered by exact row match against the token tables
    flat = out["patch_tokens"].reshape(len(tokens), -1, dim)
    src = np.full((n, 2), -1, dtype=np.int32)
    for i in range(n):
        for f in range(len(tokens)):
            hit = np.nonzero((flat[f] == feat[i]).all(axis=1))[0]
            if len(hit):
                src[i] = (f, hit[0])
                break
    assert (src >= 0).all()
    out["feat_src"] = src                                                  # (n,2): frame, patch linear index ppy*W'+ppx
    # query: image prompt through the reference's own transform + pooling + per-voxel max + sort (K=100)
    rh.flush_trimmed(m)
    from PIL import Image
    img = Image.fromarray(np.random.default_rng(3).integers(0, 256, size=(90, 120, 3), dtype=np.uint8))
    with contextlib.redirect_stdout(io.StringIO()):
        best, pos, sim = m.voxel_localized(img, K=100)
    out["query_img"] = np.asarray(img)
    out["query_vec"] = vit_ref.gaussian_pool(vit.forward_features(m.transform(img).unsqueeze(0))["x_norm_patchtokens"]).numpy()
    out["topk_pos"], out["topk_sim"], out["best_pos"] = pos, sim, best
    with contextlib.redirect_stdout(io.StringIO()):
        _, pos_r, sim_r = m.voxel_localized(img, K=30, region_radius=12, curr_grid=[int(v) for v in pos[0]])
    out["topk_pos_radius"], out["topk_sim_radius"] = pos_r, sim_r
    # cluster centres (objnav_benchmark.py:477-495) on the reference's own top-K
    robot = rh.load_agent_helpers()
    centers, labels, sizes = robot.weighted_cluster_centers(None, pos, sim, eps=10, min_samples=5)
    out["cluster_centers"], out["cluster_labels"] = np.asarray(centers, dtype=np.float64).reshape(-1, 3), labels
    return out


def run_text_goal(name="cfg1"):
    """The text-goal branch of voxel_localized (memory_2.py:566-576): `imaginary(prompt)` returns generated images (SD3.5 in the
    reference; here a stub handing back three seeded PIL images), the ViT runs on the batch of three, the pooled vectors are
    averaged.  The unmodified reference runs on config 1's store; the stub's images are kept in the fixture."""
    import types
    from PIL import Image
    ci = case_inputs(name)
    vit, dim = build_vit(ci["vit"])
    args = rh.default_args(**ci["args"])
    args.imagenary_num = 3
    rh.reset_fake_h5()
    m = rh.make_reference_memory(args, vit, patch=ci["patch"], token_dim=dim, iter_size=400000, cache_size=10 ** 9, identity_shuffle=True)
    for rgb, depth, pose in ci["frames"]:
        m.obs2voxeltoken({"rgb": rgb, "depth": depth}, pose)
    rh.flush_trimmed(m)
    imgs = [Image.fromarray(np.random.default_rng(40 + i).integers(0, 256, size=(90 + 10 * i, 120, 3), dtype=np.uint8)) for i in range(3)]
    m.imaginary = lambda prompt: types.SimpleNamespace(images=imgs)
    with contextlib.redirect_stdout(io.StringIO()):
        best, pos, sim = m.voxel_localized("a red chair next to the window", K=100)
    x = torch.stack([m.transform(im) for im in imgs])
    pooled = vit_ref.gaussian_pool(vit.forward_features(x)["x_norm_patchtokens"]).numpy()
    return {"img0": np.asarray(imgs[0]), "img1": np.asarray(imgs[1]), "img2": np.asarray(imgs[2]), "query_vec": pooled,
            "topk_pos": pos, "topk_sim": sim, "best_pos": best}


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["text_goal"]:
        out = run_text_goal()
        path = os.path.join(OUT, "cfg1_text.npz")
        np.savez_compressed(path, **out)
        print("cfg1_text", {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")
        return
    for name in sys.argv[1:] or ("cfg1", "two_pose", "sampled", "tilted"):
        out = run_reference(name)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
