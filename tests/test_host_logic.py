"""CPU: host-side logic of the product package (pose chain, config mapping) against the oracle restatement."""
import importlib

import numpy as np

from oracle import memory_port as mp


def test_pose_chain_matches_oracle_and_batch_is_bit_identical():
    geo = importlib.import_module("bsc-nav_b200.geometry")
    p = mp.MemoryParams(height=48, width=64)
    poses = np.stack([mp.circle_pose(f, 17) + np.array([0.1 * f, 0.02 * f, -0.3, 0, 0, 0, 0]) for f in range(9)])
    chain = geo.PoseChain(p.base_forward_axis, p.base_left_axis, p.base_up_axis, p.base2cam_rot, p.sensor_height)
    inv0 = None
    singles = []
    for po in poses:
        T, inv0 = mp.pose_chain(p, po, inv0)
        Tg = chain(po)
        assert np.array_equal(T, Tg)
        singles.append(Tg)
    chain2 = geo.PoseChain(p.base_forward_axis, p.base_left_axis, p.base_up_axis, p.base2cam_rot, p.sensor_height)
    assert np.array_equal(chain2.batch(poses), np.stack(singles))
    assert np.array_equal(chain2.batch(poses[3:5]), np.stack(singles[3:5]))   # T0 stays the first pose ever seen


def test_pose_chain_general_rotations():
    """Rolled / tilted poses (all nine rotation entries non-zero), not only the yaw-only poses of the synthetic circle."""
    geo = importlib.import_module("bsc-nav_b200.geometry")
    p = mp.MemoryParams(height=48, width=64)
    rng = np.random.default_rng(5)
    q = rng.normal(size=(12, 4))
    poses = np.concatenate([rng.uniform(-4, 4, size=(12, 3)), q / np.linalg.norm(q, axis=1, keepdims=True)], axis=1)
    chain = geo.PoseChain(p.base_forward_axis, p.base_left_axis, p.base_up_axis, p.base2cam_rot, p.sensor_height)
    inv0, singles = None, []
    for po in poses:
        T, inv0 = mp.pose_chain(p, po, inv0)
        Tg = chain(po)
        assert np.array_equal(T, Tg)
        singles.append(Tg)
    chain2 = geo.PoseChain(p.base_forward_axis, p.base_left_axis, p.base_up_axis, p.base2cam_rot, p.sensor_height)
    assert np.array_equal(chain2.batch(poses), np.stack(singles))


def test_geom_struct_matches_oracle_params():
    store = importlib.import_module("bsc-nav_b200.store")
    for kw in (dict(height=480, width=640, grid_size=512), dict(height=680, width=680, grid_size=1000),
               dict(height=128, width=128, grid_size=64, floor_height=-2.0, map_height=2.0)):
        p = mp.MemoryParams(**kw)
        g = store.StoreConfig(**kw).geom()
        assert (g.minh, g.maxh, g.levels, g.gs) == (p.minh, p.maxh, p.levels, p.gs)
        assert np.array_equal(np.array(g.kimg[:]).reshape(3, 3), p.calib_mat)
        assert np.array_equal(np.array(g.kinv[:]).reshape(3, 3), np.linalg.inv(p.calib_mat))
        assert np.array_equal(np.array(g.kpatch[:]).reshape(3, 3), p.patch_intr)


def test_dbscan_labels_match_sklearn():
    """Host DBSCAN labelling (floor detection, memory_2.py:205-207) == scikit-learn, the reference's dependency."""
    import importlib

    import numpy as np
    from sklearn.cluster import DBSCAN
    io = importlib.import_module("bsc-nav_b200.io")
    rng = np.random.default_rng(0)
    for _ in range(100):
        pts = np.concatenate([rng.normal(c, 0.15, size=rng.integers(0, 12)) for c in (0.0, 3.1, 6.0)]).reshape(-1, 1)
        if len(pts) == 0:
            continue
        ms = max(len(pts) // 5, 1)
        assert np.array_equal(io.dbscan_labels(pts, 0.4, ms), DBSCAN(eps=0.4, min_samples=ms).fit(pts).labels_)
    for _ in range(30):
        pts = rng.integers(0, 60, size=(80, 3)).astype(float)
        assert np.array_equal(io.dbscan_labels(pts, 10, 5), DBSCAN(eps=10, min_samples=5).fit(pts).labels_)


def test_detect_floor_ranges():
    """Two floors 3 m apart, agent on the upper one: the upper floor's cell range (memory_2.py:222-241)."""
    import importlib
    import types

    import numpy as np
    io = importlib.import_module("bsc-nav_b200.io")
    pos = np.array([[1, 1, h] for h in range(95, 160)], dtype=np.int32)
    agent = types.SimpleNamespace(get_state=lambda: types.SimpleNamespace(position=np.array([0.0, 3.05, 0.0])))
    mem = types.SimpleNamespace(base_height=[0.0, 0.02, 0.01, 3.0, 3.02, 3.05, 0.03, 2.98, 0.0, 3.01], cs=0.1, grid_rgb_pos=pos,
                                Env=types.SimpleNamespace(agent=agent), colour=False, floor_min_height=None, floor_max_height=None)
    io.detect_floor(mem)
    assert mem.num_floors == 2 and abs(mem.floor_heights[0] - 0.012) < 1e-9 and abs(mem.floor_heights[1] - 3.012) < 1e-9
    lo = 95 + (mem.floor_heights[1] - mem.floor_heights[0]) / 0.1
    assert (mem.floor_min_height, mem.floor_max_height) == (int(lo) + 1, 159 - 1)


def test_facade_grid_helpers():
    """grid2loc_2d / loc2grid_2d / _out_of_range / in_bounds of the memory class (memory_2.py:755-756, 1148-1170): plain
    arithmetic on the ctor state, exercised without a device through a stand-in `self`."""
    import types
    M = importlib.import_module("bsc-nav_b200.memory").VoxelTokenMemory
    env = types.SimpleNamespace(original_state=types.SimpleNamespace(position=np.array([1.5, 0.25, -2.0])))
    me = types.SimpleNamespace(gs=1000, cs=0.05, minh=-40, maxh=60, Env=env)
    loc = M.grid2loc_2d(me, 520, 480)                                     # (row, col)
    assert np.allclose(loc, [1.5 + (480 - 500) * 0.05, 0.25, -2.0 + (520 - 500) * 0.05])
    assert M.loc2grid_2d(me, 1.0, -2.0) == (480, 540)                     # gs/2 - int(x/cs), gs/2 - int(y/cs)
    assert M.loc2grid_2d(me, 0.049, -0.049) == (500, 500)                 # int() truncates towards zero on both sides
    assert not M._out_of_range(me, 0, 999, -40) and not M._out_of_range(me, 999, 0, 59)
    for bad in ((1000, 0, 0), (0, 1000, 0), (-1, 0, 0), (0, -1, 0), (0, 0, 60), (0, 0, -41)):
        assert M._out_of_range(me, *bad)
    assert M.in_bounds(me, 0, 999) and not M.in_bounds(me, 1000, 0) and not M.in_bounds(me, 0, -1)
