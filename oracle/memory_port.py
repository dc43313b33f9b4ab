"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's spatial-memory hot path.

This is the oracle the CUDA path is checked against.  It is pinned (tests/test_oracle_golden.py, and the
fixtures under tests/golden/ produced by oracle/make_golden.py) against the reference's own functions executed
unmodified through oracle/ref_harness.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

Each function cites the reference lines it restates.  Float64 operations use the same numpy calls, in the same
order, as the reference so that integer results (cell indices, patch indices) and the float32 winner key are
bit-identical on the same host.

Deterministic store contract: per voxel keep the C tokens with the smallest float32 r^2
(= largest alpha), ties -> lowest (frame, visiting-order) ordinal.  With <= C tokens per voxel this is the
unmodified reference store; above C the reference evicts with an unseeded random.choice (memory_2.py:352).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------- geometry (utils.py:133-214)
def cvt_pose_vec2tf(pos_quat_vec: np.ndarray) -> np.ndarray:
    """utils.py:133-141.  (px,py,pz,qx,qy,qz,qw) -> 4x4; rotation = SciPy Rotation.from_quat(...).as_matrix(),
    whose published algorithm is: normalise q, then the standard unit-quaternion matrix."""
    v = np.asarray(pos_quat_vec, dtype=np.float64).flatten()
    tf = np.eye(4)
    tf[:3, 3] = v[:3]
    q = v[3:7]
    q = q / np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    tf[0, 0] = x2 - y2 - z2 + w2
    tf[1, 0] = 2 * (xy + zw)
    tf[2, 0] = 2 * (xz - yw)
    tf[0, 1] = 2 * (xy - zw)
    tf[1, 1] = -x2 + y2 - z2 + w2
    tf[2, 1] = 2 * (yz + xw)
    tf[0, 2] = 2 * (xz + yw)
    tf[1, 2] = 2 * (yz - xw)
    tf[2, 2] = -x2 - y2 + z2 + w2
    return tf


def get_sim_cam_mat(h, w) -> np.ndarray:
    """utils.py:144-150 (used for the patch grid: fx=fy=cx=w/2, cy=h/2)."""
    k = np.eye(3)
    k[0, 0] = k[1, 1] = w / 2.0
    k[0, 2] = w / 2.0
    k[1, 2] = h / 2.0
    return k


def get_sim_cam_mat_with_fov(h, w, fov) -> np.ndarray:
    """utils.py:181-186."""
    k = np.eye(3)
    k[0, 0] = k[1, 1] = w / (2.0 * np.tan(np.deg2rad(fov / 2)))
    k[0, 2] = w / 2.0
    k[1, 2] = h / 2.0
    return k


def depth2pc(depth: np.ndarray, cam_mat: np.ndarray, min_depth=0.1, max_depth=10):
    """utils.py:153-178: camera-frame points (3,N) f64 in raster order idx = v*W+u, strict-inequality mask."""
    h, w = depth.shape
    cam_mat_inv = np.linalg.inv(cam_mat)
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    x = x.reshape((1, -1))[:, :] + 0.5
    y = y.reshape((1, -1))[:, :] + 0.5
    z = depth.reshape((1, -1))[:, :]
    p_2d = np.vstack([x, y, np.ones_like(x)])
    pc = cam_mat_inv @ p_2d
    pc = pc * z
    mask = pc[2, :] > min_depth
    mask = np.logical_and(mask, pc[2, :] < max_depth)
    return pc, mask


def sample_order(n_pix: int, rate: int, perm: np.ndarray | None = None) -> np.ndarray:
    """memory_2.py:746-749: visiting order = permutation[::rate].  perm=None -> identity (dense benchmark mode);
    to mirror the reference pass the array np.random.shuffle produced under the same seed."""
    order = np.arange(n_pix) if perm is None else np.asarray(perm)
    return order[::rate]


def backproject_depth(depth, cam_mat, min_depth, max_depth, rate=1, perm=None):
    """memory_2.py:744-753.  Returns (pc (3,N') f64, pixel index (N',)) in visiting order, mask applied after sampling."""
    pc, mask = depth2pc(depth, cam_mat, min_depth, max_depth)
    order = sample_order(pc.shape[1], rate, perm)
    m = mask[order]
    pc = pc[:, order]
    return pc[:, m], order[m]


def transform_pc(pc, pose):
    """utils.py:189-199."""
    pc_homo = np.vstack([pc, np.ones((1, pc.shape[1]))])
    return (pose @ pc_homo)[:3, :]


def base_pos2grid_id_3d(gs, cs, x_base, y_base, z_base):
    """utils.py:201-205: double truncation toward zero."""
    row = int(gs / 2 - int(x_base / cs))
    col = int(gs / 2 - int(y_base / cs))
    h = int(z_base / cs)
    return [row, col, h]


def project_point(cam_mat, p):
    """utils.py:208-214."""
    new_p = cam_mat @ p.reshape((3, 1))
    z = new_p[2, 0]
    new_p = new_p / new_p[2, 0]
    x = int(new_p[0, 0] - 0.5)
    y = int(new_p[1, 0] - 0.5)
    return x, y, z


# ----------------------------------------------------------------------------- memory state (memory_2.py:78-137, 708-722)
@dataclass
class MemoryParams:
    height: int
    width: int
    grid_size: int = 1000
    cell_size: float = 0.1
    floor_height: float = -10.0
    map_height: float = 10.0
    min_depth: float = 0.1
    max_depth: float = 10
    depth_sample_rate: int = 1
    sensor_height: float = 1.5
    n_patch_h: int = 16
    n_patch_w: int = 16
    token_dim: int = 1024
    cache_size: int = 10          # C
    bev: bool = False             # reduce the key over h (north-star 512^2 BEV grid): key (row, col), h = winner's h
    base_forward_axis: tuple = (0, 0, -1)
    base_left_axis: tuple = (-1, 0, 0)
    base_up_axis: tuple = (0, 1, 0)
    base2cam_rot: tuple = (1, 0, 0, 0, -1, 0, 0, 0, -1)
    calib_mat: np.ndarray | None = None

    def __post_init__(self):
        self.gs = self.grid_size
        self.cs = self.cell_size
        self.maxh = int(self.map_height / self.cs)      # memory_2.py:122
        self.minh = int(self.floor_height / self.cs)    # memory_2.py:123
        self.levels = self.maxh - self.minh
        if self.calib_mat is None:
            self.calib_mat = get_sim_cam_mat_with_fov(self.height, self.width, fov=90)  # memory_2.py:102
        self.base_transform = np.eye(4)                 # memory_2.py:85-88
        self.base_transform[0, :3] = self.base_forward_axis
        self.base_transform[1, :3] = self.base_left_axis
        self.base_transform[2, :3] = self.base_up_axis
        self.base2cam_tf = np.eye(4)                    # memory_2.py:90-92
        self.base2cam_tf[:3, :3] = np.array([self.base2cam_rot]).reshape((3, 3))
        self.base2cam_tf[1, 3] = self.sensor_height
        self.patch_intr = get_sim_cam_mat(self.n_patch_w, self.n_patch_h)  # memory_2.py:857: get_sim_cam_mat(shape[0], shape[1]) of the (W', H', D) array


def pose_chain(p: MemoryParams, pose, inv_init_base_tf=None):
    """memory_2.py:844-851,860.  Returns (pc_transform 4x4, inv_init_base_tf)."""
    bt = p.base_transform
    if inv_init_base_tf is None or len(inv_init_base_tf) == 0:
        init_base_tf = cvt_pose_vec2tf(pose)
        init_base_tf = bt @ init_base_tf @ np.linalg.inv(bt)
        inv_init_base_tf = np.linalg.inv(init_base_tf)
    habitat_base_pose = cvt_pose_vec2tf(pose)
    base_pose = bt @ habitat_base_pose @ np.linalg.inv(bt)
    tf = inv_init_base_tf @ base_pose
    pc_transform = tf @ bt @ p.base2cam_tf
    return pc_transform, inv_init_base_tf


def _trunc(a):
    """int() on float64: truncation toward zero (vectorised)."""
    return np.trunc(a).astype(np.int64)


def fuse_records(p: MemoryParams, depth: np.ndarray, pc_transform: np.ndarray, perm=None) -> dict:
    """Vectorised restatement of the per-point part of obs2voxeltoken (memory_2.py:859-878).

    Returns, for the points that pass BOTH the range test (:865) and the patch-index test (:878), in visiting order:
      pix   raster pixel index v*W+u of the source depth sample
      row, col, h   voxel index (h already offset by -minh, :867)
      ppx, ppy      patch-token index (token = patch_tokens[ppy, ppx], :883)
      ipx, ipy      image pixel index used for the colour lookup (rgb[ipy, ipx], :869-870; not bounds-checked there)
      r2            float64 radial distance^2 (:873); r2f = float32 value the reference stores (:885)
      alpha         exp(-r2 / 1.2) (:874-875)
    plus 'n_sampled' (points visited) and 'n_in_range'.
    """
    pc, pix = backproject_depth(depth, p.calib_mat, p.min_depth, p.max_depth, p.depth_sample_rate, perm)
    pg = transform_pc(pc, pc_transform)
    row = _trunc(p.gs / 2 - _trunc(pg[0] / p.cs))
    col = _trunc(p.gs / 2 - _trunc(pg[1] / p.cs))
    h = _trunc(pg[2] / p.cs)
    in_range = ~((col >= p.gs) | (row >= p.gs) | (h >= p.maxh) | (col < 0) | (row < 0) | (h < p.minh))  # :755-756
    # project_point with the image intrinsics and with the patch intrinsics (utils.py:208-214)
    qi = p.calib_mat @ pc
    ipx = _trunc(qi[0] / qi[2] - 0.5)
    ipy = _trunc(qi[1] / qi[2] - 0.5)
    qp = p.patch_intr @ pc
    ppx = _trunc(qp[0] / qp[2] - 0.5)
    ppy = _trunc(qp[1] / qp[2] - 0.5)
    patch_ok = ~((ppx < 0) | (ppy < 0) | (ppx >= p.n_patch_h) | (ppy >= p.n_patch_w))  # :878 (token array is (W', H', D))
    r2 = (np.square(pc[0]) + np.square(pc[1])) + np.square(pc[2])   # np.sum(np.square(p_local)), 3 terms left to right
    keep = in_range & patch_ok
    out = dict(pix=pix[keep], row=row[keep], col=col[keep], h=(h - p.minh)[keep], ppx=ppx[keep], ppy=ppy[keep],
               ipx=ipx[keep], ipy=ipy[keep], r2=r2[keep], r2f=r2[keep].astype(np.float32),
               alpha=np.exp(-r2[keep] / (2 * 0.6)), n_sampled=int(pc.shape[1]), n_in_range=int(in_range.sum()))
    return out


def fuse_frame_loop(p: MemoryParams, rgb, depth, pc_transform, patch_tokens, state, perm=None, max_points=None, stride=1):
    """Literal per-point restatement of the loop in obs2voxeltoken (memory_2.py:863-903) -- python loop, small cases
    and CPU-baseline timing only.  `state` is an OracleMemory; tokens are appended to state.pending (the 50 000-row
    host buffer of the reference without its flush-drop quirk, harness rule SURVEY 8c).  Returns points visited.
    `stride` > 1 visits every stride-th point of the visiting order (bounded, image-wide sample for CPU-baseline timing)."""
    rgb = np.array(rgb[:, :, :3])                                  # memory_2.py:853
    pc, pix = backproject_depth(depth, p.calib_mat, p.min_depth, p.max_depth, p.depth_sample_rate, perm)
    pc_global = transform_pc(pc, pc_transform)
    n = 0
    for i, (pt, p_local) in enumerate(zip(pc_global.T[::stride], pc.T[::stride])):
        if max_points is not None and i >= max_points:
            break
        n += 1
        row, col, height = base_pos2grid_id_3d(p.gs, p.cs, pt[0], pt[1], pt[2])
        if col >= p.gs or row >= p.gs or height >= p.maxh or col < 0 or row < 0 or height < p.minh:
            continue
        height = height - p.minh
        px, py, _ = project_point(p.calib_mat, p_local)
        rgb_v = rgb[py, px, :]
        px, py, _ = project_point(p.patch_intr, p_local)
        radial_dist_sq = np.sum(np.square(p_local))
        alpha = np.exp(-radial_dist_sq / (2 * 0.6))
        if not (px < 0 or py < 0 or px >= patch_tokens.shape[1] or py >= patch_tokens.shape[0]):
            state.append_token((row, col, height), np.asarray(patch_tokens[py, px, :], dtype=np.float32),
                               np.float32(radial_dist_sq))
            state.touch((row, col, height), rgb_v, alpha)
    return n


# ----------------------------------------------------------------------------- store + occupancy + colour
@dataclass
class OracleMemory:
    """Host-side model of the memory state: token store under the deterministic store rule, the occupancy index
    (memory_2.py:888-894), the alpha-weighted colour mean with per-update uint8 cast (:895-899) and the BEV
    max-height colour map (:901-903)."""
    p: MemoryParams
    tokens: dict = field(default_factory=dict)      # key -> list of (r2f float32, ordinal int, feature f32[D] | (frame, ppy, ppx))
    ordinal: int = 0
    occupied: dict = field(default_factory=dict)    # (row,col,h) -> slot, insertion (first-touch) order
    grid_rgb: list = field(default_factory=list)    # slot -> uint8[3]
    weight: list = field(default_factory=list)      # slot -> float32
    cv_map: np.ndarray | None = None
    max_height: np.ndarray | None = None

    def __post_init__(self):
        self.cv_map = np.zeros((self.p.gs, self.p.gs, 3), dtype=np.uint8)
        self.max_height = np.full((self.p.gs, self.p.gs), -np.inf)

    def _key(self, key3):
        return (key3[0], key3[1]) if self.p.bev else tuple(key3)

    def append_token(self, key3, feat, r2f, ordinal=None):
        if ordinal is None:
            ordinal = self.ordinal
            self.ordinal += 1
        self.tokens.setdefault(self._key(key3), []).append((np.float32(r2f), int(ordinal), feat, int(key3[2])))

    def touch(self, key3, rgb_v, alpha):
        """memory_2.py:888-903, literal."""
        row, col, height = key3
        slot = self.occupied.get(tuple(key3))
        if slot is None:
            self.occupied[tuple(key3)] = len(self.grid_rgb)
            self.grid_rgb.append(np.array(rgb_v, dtype=np.uint8))
            self.weight.append(np.float32(0) + np.float32(alpha))
        else:
            w = self.weight[slot]
            # numpy promotion as in the reference: u8 array * f32 scalar -> f32 array; + (u8 array * f64 scalar) -> f64; / (f32 + f64)
            mixed = (self.grid_rgb[slot] * w + rgb_v * alpha) / (w + alpha)
            self.grid_rgb[slot] = mixed.astype(np.uint8)
            self.weight[slot] = np.float32(w + alpha)
        if height >= self.max_height[row, col]:
            self.max_height[row, col] = height
            self.cv_map[row, col] = rgb_v

    def fuse_frame(self, rgb, depth, pc_transform, patch_tokens, frame_id=0, perm=None, colour=True):
        """Vectorised fusion of one frame under the store rule.  Token payloads are kept as references
        (frame_id, ppy, ppx) into `patch_tokens` to stay small; resolve with kept_tokens()."""
        p = self.p
        rec = fuse_records(p, depth, pc_transform, perm)
        hw = p.height * p.width
        for i in range(len(rec["pix"])):
            key3 = (int(rec["row"][i]), int(rec["col"][i]), int(rec["h"][i]))
            # ordinal: frame-major, then visiting order (== raster order in dense mode)
            self.append_token(key3, (frame_id, int(rec["ppy"][i]), int(rec["ppx"][i])), rec["r2f"][i],
                              ordinal=frame_id * hw + (int(rec["pix"][i]) if perm is None else i))
            if colour:
                iy, ix = int(rec["ipy"][i]), int(rec["ipx"][i])
                self.touch(key3, rgb[iy, ix, :3], rec["alpha"][i])
        return rec

    def kept(self) -> dict:
        """key -> list of the C best (r2f, ordinal, payload, h), best first."""
        C = self.p.cache_size
        return {k: sorted(v, key=lambda t: (t[0], t[1]))[:C] for k, v in self.tokens.items()}


def kept_vectorised(p: MemoryParams, frames: list) -> dict:
    """The same result as OracleMemory.fuse_frame(...) over `frames` followed by kept(), without the per-point Python loop
    (full-size parity tests: 32 frames x 307 200 points).  `frames` = [(frame_id, fuse_records(...) dict)], dense raster order
    (ordinal = frame_id * H*W + pix, memory_port.OracleMemory.fuse_frame).  Store rule as in the module docstring: per cell the C
    smallest (float32 r^2, ordinal).  Returns {cell key: [(r2f, ordinal, (frame_id, ppy, ppx), h), ...] best first}
    (tests/test_oracle_golden.py pins it against OracleMemory on small cases)."""
    hw = p.height * p.width
    cell, key, fid, ppy, ppx, hh = [], [], [], [], [], []
    for frame_id, rec in frames:
        c = rec["row"] * p.gs + rec["col"]
        cell.append(c if p.bev else c * p.levels + rec["h"])
        ordinal = np.uint64(frame_id * hw) + rec["pix"].astype(np.uint64)
        key.append((rec["r2f"].view(np.uint32).astype(np.uint64) << np.uint64(32)) | ordinal)
        fid.append(np.full(len(rec["pix"]), frame_id, dtype=np.int64))
        ppy.append(rec["ppy"]); ppx.append(rec["ppx"]); hh.append(rec["h"])
    cell, key, fid, ppy, ppx, hh = (np.concatenate(a) for a in (cell, key, fid, ppy, ppx, hh))
    order = np.lexsort((key, cell))
    cs = cell[order]
    start = np.flatnonzero(np.concatenate(([True], cs[1:] != cs[:-1])))
    rank = np.arange(len(cs)) - np.repeat(start, np.diff(np.concatenate((start, [len(cs)]))))
    sel = order[rank < p.cache_size]
    out = {}
    for i in sel:
        c = int(cell[i])
        if p.bev:
            k = (c // p.gs, c % p.gs)
        else:
            rc = c // p.levels
            k = (rc // p.gs, rc % p.gs, c % p.levels)
        r2f = np.array([int(key[i]) >> 32], dtype=np.uint32).view(np.float32)[0]
        out.setdefault(k, []).append((r2f, int(key[i]) & 0xFFFFFFFF, (int(fid[i]), int(ppy[i]), int(ppx[i])), int(hh[i])))
    return out


def cosine_similarity(q: np.ndarray, feats: np.ndarray, eps=1e-8) -> np.ndarray:
    """F.cosine_similarity(q (1,D), feats (M,D), dim=1) (memory_2.py:656), float32: w12 / sqrt(clamp(w1*w2, eps^2))."""
    q = q.astype(np.float32).reshape(1, -1)
    f = feats.astype(np.float32)
    w12 = (q * f).sum(axis=1, dtype=np.float32)
    w1 = (q * q).sum(axis=1, dtype=np.float32)
    w2 = (f * f).sum(axis=1, dtype=np.float32)
    return w12 / np.sqrt(np.maximum(w1 * w2, np.float32(eps * eps)))


def query_topk(keys3: np.ndarray, slot_feats: list, q: np.ndarray, K=100, region_radius=np.inf, curr_grid=None,
               floor_range=None):
    """memory_2.py:623-671 over an in-memory store.

    keys3 (V,3) int voxel positions [row, col, h-minh]; slot_feats[v] = (M_v, D) stored tokens of voxel v.
    Candidate filter (:624-640), cosine (:656), per-voxel max (:659-663), descending sort, first K (:665-667).
    Tie-break (contract): equal similarity -> lower (row, col, h) lexicographic.
    Returns (best (1,3) int64, positions (K',3) int64, sims (K',) float64).
    """
    keys3 = np.asarray(keys3, dtype=np.int64)
    cand = np.ones(len(keys3), dtype=bool)
    if region_radius != np.inf:
        d2 = ((keys3 - np.asarray(curr_grid, dtype=np.int64)[None, :]) ** 2).sum(axis=1)
        cand &= d2 <= region_radius ** 2
    if floor_range is not None:
        cand &= (keys3[:, 2] >= floor_range[0]) & (keys3[:, 2] <= floor_range[1])
    idx = np.nonzero(cand)[0]
    sims = np.array([cosine_similarity(q, slot_feats[v]).max() for v in idx], dtype=np.float32)
    order = sorted(range(len(idx)), key=lambda j: (-float(sims[j]), tuple(keys3[idx[j]])))[:K]
    pos = keys3[idx[order]].reshape(-1, 3)
    return pos[0:1].copy(), pos, sims[order].astype(np.float64)


def weighted_cluster_centers(top_k_positions, top_k_similarity, eps=10, min_samples=5):
    """objnav_benchmark.py:477-495 (DBSCAN is scikit-learn's, a dependency of the reference: requirements.txt)."""
    from sklearn.cluster import DBSCAN

    labels = DBSCAN(eps=eps, min_samples=min_samples).fit_predict(top_k_positions)
    info = []
    for lbl in [l for l in set(labels) if l != -1]:
        m = labels == lbl
        info.append((np.mean(top_k_similarity[m]), np.average(top_k_positions[m], axis=0, weights=top_k_similarity[m]),
                     int(np.sum(m))))
    info.sort(key=lambda t: t[0], reverse=True)
    return np.array([t[1] for t in info]), labels, [t[2] for t in info]


# ----------------------------------------------------------------------------- synthetic workload
def box_room_depth(h, w, fx, fy, cx, cy, seed):
    """Analytic box room seen from the camera (x right, y down, z forward): side walls +-4 m, floor 1.5 m below,
    ceiling 1.0 m above, end wall 6 m; plus uniform(-0.02, 0.02) noise; float32, all within (0.1, 10)."""
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
    """Pose f on a circle of `radius` m, yaw tangent: (px,py,pz,qx,qy,qz,qw) with q = (0, sin(psi/2), 0, cos(psi/2))."""
    psi = 2 * math.pi * f / max(n_frames, 1)
    return np.array([radius * math.cos(psi), 0.0, radius * math.sin(psi), 0.0, math.sin(psi / 2), 0.0, math.cos(psi / 2)])


def synthetic_frame(f, h=480, w=640, n_frames=1000, seed_base=0):
    """Config-2 frame f: rgb rng(1000+f) uniform u8 (h,w,4) as habitat delivers RGBA; depth box room rng(2000+f); pose on circle."""
    rgb = np.random.default_rng(seed_base + 1000 + f).integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    depth = box_room_depth(h, w, w / 2.0, w / 2.0, w / 2.0, h / 2.0, seed_base + 2000 + f)
    return rgb, depth, circle_pose(f, n_frames)
