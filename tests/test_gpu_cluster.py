"""Device cluster centres (csrc/cluster.cu) against the reference's weighted_cluster_centers executed through the oracle
(scikit-learn DBSCAN, objnav_benchmark.py:477-495): labels and sizes identical, centres and mean similarities to 1e-12
(float64 sums in a different association), cluster order identical."""
import numpy as np
import pytest
import torch

from oracle import memory_port as mp

pytestmark = pytest.mark.gpu


def _case(rng, n, spread, blobs):
    centres = rng.integers(0, 400, size=(blobs, 3))
    pts = centres[rng.integers(0, blobs, size=n)] + rng.integers(-spread, spread + 1, size=(n, 3))
    sims = np.sort(rng.uniform(0.2, 0.9, size=n).astype(np.float32))[::-1].copy()
    return pts.astype(np.int64), sims


@pytest.mark.parametrize("n,spread,blobs,eps,min_samples", [(100, 6, 4, 10, 5), (100, 12, 3, 10, 5), (100, 40, 2, 10, 5), (37, 5, 2, 10, 5),
                                                           (256, 9, 8, 10, 5), (1000, 15, 12, 10, 5), (100, 8, 5, 7.5, 3), (100, 3, 1, 10, 5),
                                                           (4, 1, 1, 10, 5), (100, 10, 6, 10, 1)])
def test_cluster_centers_match_sklearn(pkg, n, spread, blobs, eps, min_samples):
    cfg = pkg.StoreConfig(height=96, width=128, grid_size=64, token_dim=128, cache_size=1, bev=True)
    st = pkg.SpatialStore(cfg)
    rng = np.random.default_rng(n * 7 + spread)
    Q = 5
    cases = [_case(rng, n, spread, blobs) for _ in range(Q)]
    K = 1024 if n > 256 else 256
    pos = torch.full((Q, K, 3), -1, dtype=torch.int32)
    sim = torch.full((Q, K), float("-inf"))
    for i, (p, s) in enumerate(cases):
        pos[i, :n] = torch.from_numpy(p.astype(np.int32))
        sim[i, :n] = torch.from_numpy(s)
    cnt = torch.full((Q,), n, dtype=torch.int32)
    centers, mean_sim, sizes, labels, n_c = (t.cpu().numpy() for t in st.cluster_centers(sim.cuda(), pos.cuda(), cnt.cuda(), eps=eps,
                                                                                        min_samples=min_samples, max_clusters=256))
    for i, (p, s) in enumerate(cases):
        ref_c, ref_l, ref_sz = mp.weighted_cluster_centers(p, s.astype(np.float64), eps=eps, min_samples=min_samples)
        assert np.array_equal(labels[i, :n], ref_l), f"labels differ in case {i}"
        assert (labels[i, n:] == -1).all()
        assert n_c[i] == len(ref_sz)
        assert sizes[i, :n_c[i]].tolist() == ref_sz
        if n_c[i]:
            np.testing.assert_allclose(centers[i, :n_c[i]], ref_c, rtol=1e-12, atol=1e-9)


def test_agent_weighted_cluster_centers_golden(pkg, golden):
    """The reference's own top-K -> centres on the config-1 fixture, through the facade's agent method."""
    fx = golden("cfg1")            # the reference's own outputs (oracle/make_golden.py): topk_pos/topk_sim -> cluster_centers/labels
    synth = __import__("importlib").import_module("bsc-nav_b200.synth")
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=1, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, 0), max_batch=1)
    mem = pkg.VoxelTokenMemory(synth.Args(height=96, width=128, grid_size=64, depth_sample_rate=1), preload_dino=eng, need_diffusion=False)
    robot = pkg.GESObjectNavRobot(mem)
    centers, labels, sizes = robot.weighted_cluster_centers(fx["topk_pos"], fx["topk_sim"])
    assert np.array_equal(labels, fx["cluster_labels"])
    np.testing.assert_allclose(centers, fx["cluster_centers"], rtol=1e-9, atol=1e-9)
    assert sizes == [int((fx["cluster_labels"] == l).sum()) for l in range(len(sizes))] or len(sizes) == len(fx["cluster_centers"])
