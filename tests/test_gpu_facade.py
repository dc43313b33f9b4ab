"""GPU: the reference-facing objects (VoxelTokenMemory, LocalizeAgent) on BASELINE configs[0] against the fixture the
unmodified reference produced (tests/golden/cfg1.npz): 128x128 frame, 64^3 grid, CLIP ViT-B/32, image query."""
import importlib
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import make_golden as mg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg1(pkg):
    synth = importlib.import_module("bsc-nav_b200.synth")
    ci = mg.case_inputs("cfg1")
    cfg = pkg.CONFIGS["clip-vit-b32"]
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=4)
    args = synth.Args(**ci["args"])
    mem = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=64)
    rgb, depth, pose = ci["frames"][0]
    mem.obs2voxeltoken({"rgb": rgb, "depth": depth}, pose)
    mem.update_memory_dist_base()
    return mem, eng


def test_config1_build_matches_reference(cfg1, golden):
    mem, eng = cfg1
    g = golden("cfg1")
    assert (mem.gs, mem.cs, mem.minh, mem.maxh) == (64, 0.1, -20, 20)
    assert mem.max_id == int(g["max_id"])
    assert np.array_equal(mem.grid_rgb_pos, g["grid_rgb_pos"])          # voxel indices + slot order: bit-exact
    # stored tokens: same (voxel, r^2) multiset as the reference's buffer, features within 1e-3 relative (fp16 tensor cores)
    snap = mem.store.snapshot()
    want = {}
    tok = g["patch_tokens"][0].reshape(-1, 768)
    for pos, dis, src in zip(g["feat_pos"], g["feat_dis"], g["feat_src"]):
        want.setdefault(tuple(int(v) for v in pos), []).append((np.float32(dis), int(src[1])))
    assert set(snap) == set(want)
    worst = 0.0
    for k, items in want.items():
        items.sort(key=lambda t: t[0])
        got = snap[k]
        assert [t[0] for t in got] == [t[0] for t in items]
        for (r2, _o, _h, feat), (_r, srow) in zip(got, items):
            ref = tok[srow]
            worst = max(worst, float(np.linalg.norm(feat - ref) / np.linalg.norm(ref)))
    assert worst < 1e-3, worst
    assert mem.get_total_token_count() == sum(len(v) for v in snap.values())      # memory_2.py:312-323
    assert mem.in_bounds(0, 63) and not mem.in_bounds(64, 0) and mem._out_of_range(0, 0, 20)


def test_config1_image_query_matches_reference(cfg1, golden):
    mem, eng = cfg1
    g = golden("cfg1")
    img = Image.fromarray(g["query_img"])
    q = mem._query_vector(img).cpu().numpy()
    assert np.linalg.norm(q - g["query_vec"]) / np.linalg.norm(g["query_vec"]) < 1e-3
    best, pos, sim = mem.voxel_localized(img, K=100)
    assert pos.shape == (100, 3) and pos.dtype == np.int64 and sim.dtype == np.float64 and best.shape == (1, 3)
    assert np.all(np.diff(sim) <= 0)
    np.testing.assert_allclose(sim, g["topk_sim"], rtol=1e-3, atol=1e-4)
    # position sets: exact wherever the reference's ranking is separated from the K-th similarity by more than the tolerance
    cut = g["topk_sim"][-1] + 2e-3
    sure = {tuple(r) for r, s in zip(g["topk_pos"], g["topk_sim"]) if s > cut}
    assert sure <= {tuple(r) for r in pos}


def test_localize_agent_and_roundtrip(cfg1, golden, tmp_path, pkg):
    mem, eng = cfg1
    g = golden("cfg1")
    agent = pkg.LocalizeAgent(mem, None, save_dir=str(tmp_path / "localize_results"))
    centers = agent.LocalizeImagePrompt(Image.fromarray(g["query_img"]))
    saved = np.load(tmp_path / "localize_results" / "best_pos_topK_img_input.npy")
    assert saved.shape == (100, 3) and saved.dtype == np.int64                  # format of localize_results/*.npy
    assert centers.ndim == 2 and centers.shape[1] == 3
    # save -> load into a fresh memory -> identical query result
    path = str(tmp_path / "mem")
    mem.save_memory(path)
    for f in ("grid_rgb_pos.npy", "max_id.npy", "occupied_ids.npy", "map_height.npy", "base_height.npy", "long_memory.json"):
        assert os.path.exists(os.path.join(path, f))
    synth = importlib.import_module("bsc-nav_b200.synth")
    args = synth.Args(**mg.case_inputs("cfg1")["args"], load_memory_path=path)
    mem2 = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=64)
    mem2.load_memory()
    q = torch.from_numpy(g["query_vec"])
    _, p1, s1 = mem.voxel_localized(q, K=50)
    _, p2, s2 = mem2.voxel_localized(q, K=50)
    assert np.array_equal(p1, p2) and np.array_equal(s1, s2)
    assert mem2.max_id == mem.max_id and np.array_equal(mem2.grid_rgb_pos, mem.grid_rgb_pos)


def test_text_goal_branch_matches_reference(cfg1, golden):
    """voxel_localized(str) (memory_2.py:566-576): the injected generator hands back three images, the engine runs the batch of
    three, the pooled vectors are averaged -- against the unmodified reference run with the same stub (tests/golden/cfg1_text.npz,
    oracle/make_golden.py text_goal)."""
    import types
    mem, eng = cfg1
    g = golden("cfg1_text")
    imgs = [Image.fromarray(g[f"img{i}"]) for i in range(3)]
    calls = []
    mem.imaginary = lambda prompt: (calls.append(prompt), types.SimpleNamespace(images=imgs))[1]
    mem.args.imagenary_num = 3
    q = mem._query_vector("a red chair next to the window").cpu().numpy()
    assert calls == ["a red chair next to the window"] and q.shape == (1, 768)
    assert np.linalg.norm(q - g["query_vec"]) / np.linalg.norm(g["query_vec"]) < 1e-3
    best, pos, sim = mem.voxel_localized("a red chair next to the window", K=100)
    assert pos.shape == (100, 3) and sim.dtype == np.float64
    np.testing.assert_allclose(sim, g["topk_sim"], rtol=1e-3, atol=1e-4)
    cut = g["topk_sim"][-1] + 2e-3
    sure = {tuple(r) for r, s in zip(g["topk_pos"], g["topk_sim"]) if s > cut}
    assert sure <= {tuple(r) for r in pos}
    del mem.imaginary                       # back to the class's NotImplementedError for the next user of the fixture
    with pytest.raises(NotImplementedError):
        mem.voxel_localized("no generator injected")


def test_fuse_frames_input_forms_agree(pkg):
    """fuse_frames takes numpy arrays, nested lists, other dtypes for the depth, pinned host tensors and device tensors; the depth is
    staged and copied behind the ViT (memory.py: _inputs_to_device).  Every form must build the same store, bit for bit."""
    synth = importlib.import_module("bsc-nav_b200.synth")
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=2, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=3, perturb=True), max_batch=4)
    H, W = 96, 128
    frames = [synth.synthetic_frame(f, H, W, n_frames=12) for f in range(6)]
    rgb = np.stack([f[0] for f in frames]); depth = np.stack([f[1] for f in frames]).astype(np.float32)
    poses = np.stack([synth.tour_pose(f) for f in range(6)])
    forms = {
        "numpy": lambda a, b: (rgb[a:b], depth[a:b]),
        "float64 depth": lambda a, b: (rgb[a:b], depth[a:b].astype(np.float64)),
        "lists": lambda a, b: (rgb[a:b], depth[a:b].tolist()),
        "pinned": lambda a, b: (torch.from_numpy(rgb[a:b].copy()).pin_memory(), torch.from_numpy(depth[a:b].copy()).pin_memory()),
        "device": lambda a, b: (torch.from_numpy(rgb[a:b]).cuda(), torch.from_numpy(depth[a:b]).cuda()),
        "host rgb, device depth": lambda a, b: (rgb[a:b], torch.from_numpy(depth[a:b]).cuda()),
    }
    snaps = {}
    for name, form in forms.items():
        mem = pkg.VoxelTokenMemory(synth.Args(height=H, width=W, grid_size=64, depth_sample_rate=1), preload_dino=eng,
                                   need_diffusion=False, cache_size=1, max_batch=4)      # C = 1: the row layout itself is deterministic
        keep = []
        for a, b in ((0, 3), (3, 4), (4, 6)):                 # ragged batches: the staging and device buffers are re-used and re-sized
            r, d = form(a, b)
            keep.append((r, d))                               # pinned inputs stay alive and unchanged until the copies have run
            mem.fuse_frames(r, d, poses[a:b])
        mem.update_memory_dist_base()
        st = mem.store
        n = mem.max_id
        snaps[name] = (n, st.slot_cell[:n].cpu(), st.row_key[:n].cpu(), st.tok_h[:n].cpu(), st.feat[:n].cpu())
    ref = snaps["device"]
    assert ref[0] > 0
    for name, s in snaps.items():
        assert s[0] == ref[0], name
        for x, y in zip(s[1:], ref[1:]):
            assert torch.equal(x, y), name
