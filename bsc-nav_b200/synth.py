"""Synthetic RGB-D workload of BASELINE.json configs[1..4]: analytic box room + noise,
uniform-random colour, poses on a circle.  Seeded; identical arrays to oracle/memory_port.py's generators
(tests/test_synth.py keeps the two in lock-step)."""
from __future__ import annotations

import math

import numpy as np


def box_room_depth(h, w, fx, fy, cx, cy, seed):
    v, u = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    dx = (u + 0.5 - cx) / fx
    dy = (v + 0.5 - cy) / fy
    with np.errstate(divide="ignore"):
        z = np.minimum.reduce([
            np.full((h, w), 6.0),
            np.where(dx > 0, 4.0 / np.abs(dx), np.inf), np.where(dx < 0, 4.0 / np.abs(dx), np.inf),
            np.where(dy > 0, 1.5 / np.abs(dy), np.inf), np.where(dy < 0, 1.0 / np.abs(dy), np.inf),
        ])
    z = z + np.random.default_rng(seed).uniform(-0.02, 0.02, size=(h, w))
    return z.astype(np.float32)


def circle_pose(f, n_frames, radius=1.5):
    psi = 2 * math.pi * f / max(n_frames, 1)
    return np.array([radius * math.cos(psi), 0.0, radius * math.sin(psi), 0.0, math.sin(psi / 2), 0.0, math.cos(psi / 2)])


def tour_pose(f, period=1000):
    """Pose f of a tour that never repeats: yaw advances 2 pi / period per frame, the radius breathes with an incommensurate
    period.  bench.py cycles a pool of depth/colour images (camera-frame box room) under these poses, so every step touches new
    cells and new (r^2, ordinal) winners."""
    psi = 2 * math.pi * f / period
    radius = 1.5 + 0.9 * math.sin(2 * math.pi * f / (period * 0.6180339887))
    return np.array([radius * math.cos(psi), 0.0, radius * math.sin(psi), 0.0, math.sin(psi / 2), 0.0, math.cos(psi / 2)])


def synthetic_frame(f, h=480, w=640, n_frames=1000, seed_base=0):
    rgb = np.random.default_rng(seed_base + 1000 + f).integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    depth = box_room_depth(h, w, w / 2.0, w / 2.0, w / 2.0, h / 2.0, seed_base + 2000 + f)
    return rgb, depth, circle_pose(f, n_frames)


class Args:
    """The argparse namespace fields the memory reads (args.py:24-67), with the reference's defaults."""

    def __init__(self, **over):
        d = dict(height=680, width=680, query_height=224, query_width=224, cell_size=0.1, grid_size=1000,
                 depth_sample_rate=1000, min_depth=0.1, max_depth=10, floor_height=-10.0, map_height=10.0,
                 sensor_height=1.5, base_forward_axis=[0, 0, -1], base_left_axis=[-1, 0, 0], base_up_axis=[0, 1, 0],
                 base2cam_rot=[1, 0, 0, 0, -1, 0, 0, 0, -1], load_single_floor=False, imagenary_num=3,
                 dino_size="dinov2_vitl14_reg", memory_path="./memory", scene_name="scene", load_memory_path=None,
                 random_move_num=1, turn_left=30)
        d.update(over)
        self.__dict__.update(d)
