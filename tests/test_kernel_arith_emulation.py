"""CPU: the arithmetic of k_index_points (csrc/memory_kernels.cu) restated operation by operation -- float64 with exactly
rounded fused multiply-adds in the kernel's order -- against the reference's own buffers (tests/golden/tilted.npz: rolled and
tilted poses, depth holes, points at and beyond both depth bounds).  It pins the *statement* of the kernel's index math
(the dgemm-order FMA chains, the double truncations, the strict depth bounds) without a GPU; the GPU test
(test_gpu_memory.py) then only has to show that the device executes that statement."""
import importlib
from fractions import Fraction

import numpy as np

from oracle import make_golden as mg


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))     # one rounding, like __fma_rn


def _dot3(m, a, b, c):
    return _fma(m[2], c, _fma(m[1], b, m[0] * a))


def _dot4(m, a, b, c, d):
    return _fma(m[3], d, _fma(m[2], c, _fma(m[1], b, m[0] * a)))


def _trunc(v):
    if not v > -2.0e9:
        return -2000000000
    if not v < 2.0e9:
        return 2000000000
    return int(v)


def test_index_math_statement_matches_reference_buffers(golden):
    g = golden("tilted")
    ci = mg.case_inputs("tilted")
    a = ci["args"]
    store = importlib.import_module("bsc-nav_b200.store")
    geo = importlib.import_module("bsc-nav_b200.geometry")
    geom = store.StoreConfig(height=a["height"], width=a["width"], grid_size=a["grid_size"], floor_height=a["floor_height"],
                             map_height=a["map_height"], n_patch_h=16, n_patch_w=16, token_dim=128, cache_size=1).geom()
    kinv, kpatch = list(geom.kinv), list(geom.kpatch)
    chain = geo.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
    H, W, gs, cs = geom.height, geom.width, geom.gs, geom.cs
    pos, dis, src = [], [], []
    for f, (_rgb, depth, pose) in enumerate(ci["frames"]):
        M = chain(pose).reshape(-1).tolist()
        for pix in range(H * W):
            z = float(depth.reshape(-1)[pix])
            v, u = divmod(pix, W)
            x, y = u + 0.5, v + 0.5
            p0, p1, p2 = _dot3(kinv[0:3], x, y, 1.0) * z, _dot3(kinv[3:6], x, y, 1.0) * z, _dot3(kinv[6:9], x, y, 1.0) * z
            if not (geom.min_depth < p2 < geom.max_depth):
                continue
            gx, gy, gz = _dot4(M[0:4], p0, p1, p2, 1.0), _dot4(M[4:8], p0, p1, p2, 1.0), _dot4(M[8:12], p0, p1, p2, 1.0)
            row, col, h = _trunc(gs / 2.0 - float(_trunc(gx / cs))), _trunc(gs / 2.0 - float(_trunc(gy / cs))), _trunc(gz / cs)
            if col >= gs or row >= gs or h >= geom.maxh or col < 0 or row < 0 or h < geom.minh:
                continue
            q0, q1, q2 = _dot3(kpatch[0:3], p0, p1, p2), _dot3(kpatch[3:6], p0, p1, p2), _dot3(kpatch[6:9], p0, p1, p2)
            ppx, ppy = _trunc(q0 / q2 - 0.5), _trunc(q1 / q2 - 0.5)
            if ppx < 0 or ppy < 0 or ppx >= geom.n_patch_h or ppy >= geom.n_patch_w:
                continue
            pos.append((row, col, h - geom.minh))
            dis.append(np.float32((p0 * p0 + p1 * p1) + p2 * p2))
            src.append((f, ppy * geom.n_patch_h + ppx))
    assert len(pos) == len(g["feat_pos"]) > 4000
    assert np.array_equal(np.array(pos), g["feat_pos"])                 # voxel index of every accepted point, in visiting order
    assert np.array_equal(np.array(dis, dtype=np.float32), g["feat_dis"])
    assert np.array_equal(np.array(src), g["feat_src"])


def test_guarded_multiply_equals_division_then_truncation():
    """k_index_points replaces its five IEEE divisions per point by multiplications guarded next to integers
    (trunc_div_cs / trunc_quot_half in csrc/memory_kernels.cu): the guarded forms return the integer the division returns, on
    random inputs and on inputs planted on and next to cell boundaries (the long form: tools/micro/div_trunc_check.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "micro"))
    import div_trunc_check
    total, slow = div_trunc_check.check(n_rounds=2, size=300_000)
    assert total > 4_000_000 and slow > 100_000          # the planted boundary cases do exercise the exact-division path
