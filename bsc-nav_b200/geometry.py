"""Host-side geometry of the memory: 4x4 / 3x3 float64 matrices only (the per-point math runs on the device).

Mirrors the reference's helpers so that the matrices handed to the kernels are bit-identical to the ones the
reference computes: utils.py:133-150,181-186 and the pose chain of memory_2.py:844-851,860.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as R


def cvt_pose_vec2tf(pos_quat_vec) -> np.ndarray:
    """(px,py,pz,qx,qy,qz,qw) -> 4x4 (utils.py:133-141)."""
    v = np.asarray(pos_quat_vec, dtype=np.float64)
    tf = np.eye(4)
    tf[:3, 3] = v[:3].flatten()
    tf[:3, :3] = R.from_quat(v[3:].flatten()).as_matrix()
    return tf


def sim_cam_mat(h, w) -> np.ndarray:
    """utils.py:144-150."""
    k = np.eye(3)
    k[0, 0] = k[1, 1] = w / 2.0
    k[0, 2] = w / 2.0
    k[1, 2] = h / 2.0
    return k


def sim_cam_mat_with_fov(h, w, fov) -> np.ndarray:
    """utils.py:181-186."""
    k = np.eye(3)
    k[0, 0] = k[1, 1] = w / (2.0 * np.tan(np.deg2rad(fov / 2)))
    k[0, 2] = w / 2.0
    k[1, 2] = h / 2.0
    return k


class PoseChain:
    """pc_transform = inv(T0) @ B P(pose) B^-1 @ B @ base2cam, T0 from the first pose seen (memory_2.py:844-851,860)."""

    def __init__(self, base_forward_axis, base_left_axis, base_up_axis, base2cam_rot, sensor_height):
        self.base_transform = np.eye(4)
        self.base_transform[0, :3] = base_forward_axis
        self.base_transform[1, :3] = base_left_axis
        self.base_transform[2, :3] = base_up_axis
        self.base2cam_tf = np.eye(4)
        self.base2cam_tf[:3, :3] = np.array([base2cam_rot]).reshape((3, 3))
        self.base2cam_tf[1, 3] = sensor_height
        self.inv_init_base_tf = []
        self.tf = None

    def reset(self):
        self.inv_init_base_tf = []

    def batch(self, poses) -> np.ndarray:
        """(B,7) poses -> (B,4,4) pc_transforms; same operations as __call__, stacked (one SciPy / matmul call per step)."""
        poses = np.asarray(poses, dtype=np.float64).reshape(-1, 7)
        if len(self.inv_init_base_tf) == 0:
            self(poses[0])
        bt = self.base_transform
        tf = np.tile(np.eye(4), (len(poses), 1, 1))
        tf[:, :3, 3] = poses[:, :3]
        tf[:, :3, :3] = R.from_quat(poses[:, 3:]).as_matrix()
        base_pose = bt @ tf @ np.linalg.inv(bt)
        tfs = self.inv_init_base_tf @ base_pose
        self.tf = tfs[-1]
        return tfs @ bt @ self.base2cam_tf

    def __call__(self, pose) -> np.ndarray:
        bt = self.base_transform
        if len(self.inv_init_base_tf) == 0:
            init_base_tf = cvt_pose_vec2tf(pose)
            init_base_tf = bt @ init_base_tf @ np.linalg.inv(bt)
            self.inv_init_base_tf = np.linalg.inv(init_base_tf)
        habitat_base_pose = cvt_pose_vec2tf(pose)
        base_pose = bt @ habitat_base_pose @ np.linalg.inv(bt)
        self.tf = self.inv_init_base_tf @ base_pose
        return self.tf @ bt @ self.base2cam_tf
