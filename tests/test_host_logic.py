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
