"""CPU: the numpy oracle (oracle/memory_port.py) against the fixtures produced by the unmodified reference
(oracle/make_golden.py), and -- when /root/reference is mounted -- against the reference executed live."""
import numpy as np
import pytest

from oracle import memory_port as mp
from oracle import make_golden as mg
from oracle import ref_harness as rh

CASES = ("cfg1", "two_pose", "sampled", "tilted")


def port_run(name, g):
    ci = mg.case_inputs(name) if isinstance(name, str) else name
    a = ci["args"]
    tok = g["patch_tokens"]
    p = mp.MemoryParams(height=a["height"], width=a["width"], grid_size=a["grid_size"],
                        floor_height=a.get("floor_height", -10.0), map_height=a.get("map_height", 10.0),
                        cell_size=a.get("cell_size", 0.1), depth_sample_rate=a["depth_sample_rate"], n_patch_h=tok.shape[2], n_patch_w=tok.shape[1],
                        token_dim=tok.shape[3], cache_size=10 ** 9)
    om = mp.OracleMemory(p)
    inv0 = None
    for f, (rgb, depth, pose) in enumerate(ci["frames"]):
        T, inv0 = mp.pose_chain(p, pose, inv0)
        perm = None
        if "sample_idx" in g:   # visiting order stored by the generator; rate already applied
            perm = g["sample_idx"][f]
            p.depth_sample_rate = 1
        om.fuse_frame(rgb, depth, T, tok[f], frame_id=f, perm=perm)
    return p, om


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference_fixture(name, golden):
    g = golden(name)
    p, om = port_run(name, g)
    items = sorted(((t[1], k, t) for k, v in om.tokens.items() for t in v), key=lambda x: x[0])
    pos = np.array([[k[0], k[1], t[3]] for _, k, t in items])
    dis = np.array([t[0] for _, k, t in items], dtype=np.float32)
    src = np.array([[t[2][0], t[2][1] * p.n_patch_w + t[2][2]] for _, k, t in items])
    assert np.array_equal(pos, g["feat_pos"])            # voxel indices: bit-exact
    assert np.array_equal(dis, g["feat_dis"])            # float32 winner key: bit-exact
    assert np.array_equal(src, g["feat_src"])            # which patch token each point stored
    assert len(om.occupied) == int(g["max_id"])
    assert np.array_equal(np.array(list(om.occupied.keys())), g["grid_rgb_pos"])   # first-touch slot order
    assert np.array_equal(np.array(om.grid_rgb), g["grid_rgb"])
    assert np.array_equal(np.array(om.weight, dtype=np.float32), g["weight"])
    assert np.array_equal(om.cv_map, g["cv_map"])
    assert np.array_equal(om.max_height, g["max_height"])
    assert (p.minh, p.maxh) == tuple(g["minh_maxh"])


@pytest.mark.parametrize("name", CASES)
def test_port_query_matches_reference_fixture(name, golden):
    g = golden(name)
    p, om = port_run(name, g)
    tok = g["patch_tokens"]
    keys = sorted(om.tokens.keys())
    feats = [np.stack([tok[t[2][0], t[2][1], t[2][2]] for t in sorted(om.tokens[k], key=lambda t: t[1])]) for k in keys]
    _, pos, sim = mp.query_topk(np.array(keys), feats, g["query_vec"], K=100)
    assert pos.dtype == np.int64 and sim.dtype == np.float64 and pos.shape == g["topk_pos"].shape
    np.testing.assert_allclose(sim, g["topk_sim"], rtol=0, atol=2e-7)
    # sets agree wherever the reference's ordering is unambiguous (strictly above the K-th similarity);
    # exact ties at the boundary are ordered by HDF5 name order there and by (row,col,h) in the contract
    cut = g["topk_sim"][-1] + 1e-6
    ref_set = {tuple(r) for r, s in zip(g["topk_pos"], g["topk_sim"]) if s > cut}
    assert ref_set <= {tuple(r) for r in pos}
    _, pos_r, sim_r = mp.query_topk(np.array(keys), feats, g["query_vec"], K=30, region_radius=12,
                                    curr_grid=[int(v) for v in g["topk_pos"][0]])
    np.testing.assert_allclose(sim_r, g["topk_sim_radius"], rtol=0, atol=2e-7)
    cut = g["topk_sim_radius"][-1] + 1e-6
    assert {tuple(r) for r, s in zip(g["topk_pos_radius"], g["topk_sim_radius"]) if s > cut} <= {tuple(r) for r in pos_r}


def test_cluster_centers_fixture(golden):
    g = golden("two_pose")
    centers, labels, _ = mp.weighted_cluster_centers(g["topk_pos"], g["topk_sim"])
    np.testing.assert_allclose(centers, g["cluster_centers"], rtol=1e-12)
    assert np.array_equal(labels, g["cluster_labels"])


def test_geometry_helpers_match_reference_live():
    if not rh.available():
        pytest.skip("/root/reference not mounted (GPU box)")
    ns = rh.load_utils_namespace()
    rng = np.random.default_rng(0)
    for _ in range(50):
        v = np.concatenate([rng.uniform(-5, 5, 3), rng.normal(size=4)])
        np.testing.assert_allclose(mp.cvt_pose_vec2tf(v), ns["cvt_pose_vec2tf"](v), rtol=0, atol=1e-15)
    assert np.array_equal(mp.get_sim_cam_mat(16, 16), ns["get_sim_cam_mat"](16, 16))
    assert np.array_equal(mp.get_sim_cam_mat_with_fov(480, 640, 90), ns["get_sim_cam_mat_with_fov"](480, 640, 90))
    d = rng.uniform(0.05, 12, size=(24, 32)).astype(np.float32)
    k = mp.get_sim_cam_mat_with_fov(24, 32, 90)
    pc, m = mp.depth2pc(d, k)
    pc2, m2 = ns["depth2pc"](d, intr_mat=k)
    assert np.array_equal(pc, pc2) and np.array_equal(m, m2)
    for x, y, z in rng.uniform(-7, 7, size=(200, 3)):
        assert mp.base_pos2grid_id_3d(100, 0.1, x, y, z) == ns["base_pos2grid_id_3d"](100, 0.1, x, y, z)
        pt = np.array([x, y, abs(z) + 0.2])
        assert mp.project_point(k, pt) == ns["project_point"](k, pt)


def test_loop_restatement_equals_vectorised(golden):
    """fuse_frame_loop (literal per-point form, used for CPU-baseline timing) == vectorised fuse_frame."""
    g = golden("two_pose")
    ci = mg.case_inputs("two_pose")
    p, om = port_run("two_pose", g)
    om2 = mp.OracleMemory(p)
    inv0 = None
    for f, (rgb, depth, pose) in enumerate(ci["frames"]):
        T, inv0 = mp.pose_chain(p, pose, inv0)
        mp.fuse_frame_loop(p, rgb, depth, T, g["patch_tokens"][f], om2)
    assert list(om.occupied.keys()) == list(om2.occupied.keys())
    assert np.array_equal(np.array(om.grid_rgb), np.array(om2.grid_rgb))
    assert sum(len(v) for v in om.tokens.values()) == sum(len(v) for v in om2.tokens.values())


def _random_case(seed):
    """Small random scene for the live cross-check: tilted / rolled poses (not only yaw), depth holes (0), points beyond
    max_depth and below min_depth, points leaving the grid and the height range, a non-default cell size."""
    rng = np.random.default_rng(seed)
    h, w = (40, 56) if seed % 2 else (48, 48)
    frames = []
    for f in range(3):
        rgb = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        depth = rng.uniform(0.3, 6.0, size=(h, w)).astype(np.float32)
        depth[rng.random((h, w)) < 0.1] = 0.0                       # sensor holes
        depth[rng.random((h, w)) < 0.05] = 25.0                     # beyond max_depth
        depth[rng.random((h, w)) < 0.03] = 0.1                      # exactly min_depth: strict bound, dropped (utils.py:174-176)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if f == 0:
            q = np.array([0.0, 0.0, 0.0, 1.0])
        pose = np.concatenate([rng.uniform(-1.5, 1.5, 3) * (f > 0), q])
        frames.append((rgb, depth, pose))
    return dict(args=dict(height=h, width=w, grid_size=40, depth_sample_rate=1 if seed % 3 else 7, cell_size=[0.1, 0.07, 0.25][seed % 3],
                          floor_height=-1.0, map_height=1.5), frames=frames, patch=14, vit=("table", 256, 16, 11 + seed),
                **({} if seed % 3 else {"shuffle_seed": 3 + seed}))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_port_matches_reference_live_random(seed):
    """Beyond the three committed fixtures: the port against the unmodified reference executed here on random scenes."""
    if not rh.available():
        pytest.skip("/root/reference not mounted (GPU box)")
    ci = _random_case(seed)
    g = mg.run_reference(ci)
    p, om = port_run(ci, g)
    items = sorted(((t[1], k, t) for k, v in om.tokens.items() for t in v), key=lambda x: x[0])
    assert len(items) == len(g["feat_dis"]) > 50
    assert np.array_equal(np.array([[k[0], k[1], t[3]] for _, k, t in items]), g["feat_pos"])
    assert np.array_equal(np.array([t[0] for _, k, t in items], dtype=np.float32), g["feat_dis"])
    assert np.array_equal(np.array([[t[2][0], t[2][1] * p.n_patch_w + t[2][2]] for _, k, t in items]), g["feat_src"])
    assert np.array_equal(np.array(list(om.occupied.keys())), g["grid_rgb_pos"])
    assert np.array_equal(np.array(om.grid_rgb), g["grid_rgb"])
    assert np.array_equal(np.array(om.weight, dtype=np.float32), g["weight"])
    assert np.array_equal(om.cv_map, g["cv_map"]) and np.array_equal(om.max_height, g["max_height"])
    # and the query over it
    tok = g["patch_tokens"]
    keys = sorted(om.tokens.keys())
    feats = [np.stack([tok[t[2][0], t[2][1], t[2][2]] for t in sorted(om.tokens[k], key=lambda t: t[1])]) for k in keys]
    _, pos, sim = mp.query_topk(np.array(keys), feats, g["query_vec"], K=100)
    np.testing.assert_allclose(sim, g["topk_sim"], rtol=0, atol=2e-7)
    cut = g["topk_sim"][-1] + 1e-6
    assert {tuple(r) for r, s in zip(g["topk_pos"], g["topk_sim"]) if s > cut} <= {tuple(r) for r in pos}


@pytest.mark.parametrize("bev,C", [(True, 1), (False, 1), (False, 3), (True, 2)])
def test_kept_vectorised_equals_oracle_memory(golden, bev, C):
    """kept_vectorised (full-size GPU parity tests) == OracleMemory.fuse_frame + kept() on the two-pose fixture's inputs."""
    ci = mg.case_inputs("two_pose")
    a = ci["args"]
    p = mp.MemoryParams(height=a["height"], width=a["width"], grid_size=a["grid_size"], floor_height=a.get("floor_height", -10.0),
                        map_height=a.get("map_height", 10.0), depth_sample_rate=1, token_dim=8, cache_size=C, bev=bev)
    om = mp.OracleMemory(p)
    inv0, frames = None, []
    for f, (rgb, depth, pose) in enumerate(ci["frames"]):
        T, inv0 = mp.pose_chain(p, pose, inv0)
        frames.append((f, om.fuse_frame(rgb, depth, T, None, frame_id=f, colour=False)))
    want, got = om.kept(), mp.kept_vectorised(p, frames)
    assert set(want) == set(got)
    for k, items in want.items():
        assert [(t[0], t[1], t[2], t[3]) for t in items] == got[k]


def test_text_goal_fixture_query_vector(golden):
    """The text-goal branch's query vector (memory_2.py:566-608, three generated images): the oracle's pooling over a batch of three
    reproduces what the unmodified reference computed (tests/golden/cfg1_text.npz)."""
    import torch
    from PIL import Image
    from torchvision import transforms
    from oracle import vit_ref
    g = golden("cfg1_text")
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    x = torch.stack([tf(Image.fromarray(g[f"img{i}"])) for i in range(3)])
    vit, _dim = mg.build_vit(mg.case_inputs("cfg1")["vit"])
    q = vit_ref.gaussian_pool(vit.forward_features(x)["x_norm_patchtokens"]).numpy()
    np.testing.assert_allclose(q, g["query_vec"], rtol=1e-5, atol=1e-6)
    assert g["topk_pos"].shape == (100, 3) and np.all(np.diff(g["topk_sim"]) <= 0)
