"""Device-resident spatial store: the HBM layout behind VoxelTokenMemory and the thin wrappers over the C ABI.

Replaces the reference's host buffers + HDF5 token file (memory_2.py:708-722, 326-358): occupancy index, per-cell
top-C winner keys, token rows and their norms all live in HBM (layout in include/bscnav_b200.h: bscnav_store_t).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .geometry import sim_cam_mat, sim_cam_mat_with_fov


@dataclass
class StoreConfig:
    height: int
    width: int
    grid_size: int = 1000
    cell_size: float = 0.1
    floor_height: float = -10.0
    map_height: float = 10.0
    min_depth: float = 0.1
    max_depth: float = 10
    n_patch_h: int = 16
    n_patch_w: int = 16
    token_dim: int = 1024
    cache_size: int = 10           # C (memory_2.py:111)
    bev: bool = False              # key reduced over h: the north-star "512x512 BEV feature grid"
    max_slots: int | None = None   # default gs*gs, the reference's capacity (memory_2.py:715-718)
    calib_mat: np.ndarray | None = None
    colour: bool = False           # keep grid_rgb / weight / cv_map / max_height (memory_2.py:888-903); voxel keys only

    def geom(self) -> _lib.Geom:
        gs, cs = int(self.grid_size), float(self.cell_size)
        maxh = int(self.map_height / cs)       # memory_2.py:122
        minh = int(self.floor_height / cs)     # memory_2.py:123
        k = self.calib_mat if self.calib_mat is not None else sim_cam_mat_with_fov(self.height, self.width, fov=90)
        k = np.asarray(k, dtype=np.float64)
        kinv = np.linalg.inv(k)                # utils.py:164
        kp = sim_cam_mat(self.n_patch_w, self.n_patch_h)   # get_sim_cam_mat(shape[0], shape[1]) of the (W', H', D) token array
        g = _lib.Geom(height=self.height, width=self.width, gs=gs, minh=minh, maxh=maxh,
                      levels=1 if self.bev else maxh - minh, n_patch_h=self.n_patch_h, n_patch_w=self.n_patch_w,
                      cache=int(self.cache_size), dim=int(self.token_dim),
                      max_slots=int(self.max_slots if self.max_slots is not None else gs * gs), reserved=0,
                      cs=cs, min_depth=float(self.min_depth), max_depth=float(self.max_depth))
        g.kinv[:] = kinv.reshape(-1).tolist()
        g.kimg[:] = k.reshape(-1).tolist()
        g.kpatch[:] = kp.reshape(-1).tolist()
        return g


class SpatialStore:
    def __init__(self, cfg: StoreConfig, device="cuda"):
        self.lib = _lib.lib()
        self.cfg = cfg
        self.device = torch.device(device)
        self.g = cfg.geom()
        g = self.g
        cells = g.gs * g.gs * g.levels
        rows = g.max_slots * g.cache
        dev = self.device
        self.occupied_ids = torch.empty(cells, dtype=torch.int32, device=dev)
        self.first_touch = torch.empty(cells, dtype=torch.int32, device=dev)
        self.slot_cell = torch.empty(g.max_slots, dtype=torch.int32, device=dev)
        self.slot_keys = torch.empty(rows, dtype=torch.int64, device=dev)
        self.row_key = torch.empty(rows, dtype=torch.int64, device=dev)
        self.tok_h = torch.empty(rows, dtype=torch.int32, device=dev)
        self.norm2 = torch.empty(rows, dtype=torch.float32, device=dev)
        self.feat = torch.empty(rows, g.dim, dtype=torch.float32, device=dev)
        self.counters = torch.zeros(8, dtype=torch.int32, device=dev)
        self.st = _lib.Store(*[C.c_void_p(t.data_ptr()) for t in (self.occupied_ids, self.first_touch, self.slot_cell,
                                                                   self.slot_keys, self.row_key, self.tok_h, self.norm2,
                                                                   self.feat, self.counters)])
        self.col = None
        if cfg.colour:
            if g.levels == 1:
                raise _lib.BscnavError("colour channels are defined per voxel (row, col, h): not available with bev=True")
            self.grid_rgb = torch.empty(g.max_slots, 3, dtype=torch.uint8, device=dev)
            self.weight = torch.empty(g.max_slots, dtype=torch.float32, device=dev)
            self.cv_key = torch.empty(g.gs * g.gs, dtype=torch.int64, device=dev)
            self.cv_map = torch.empty(g.gs, g.gs, 3, dtype=torch.uint8, device=dev)
            self.col = _lib.Colour(*[C.c_void_p(t.data_ptr()) for t in (self.grid_rgb, self.weight, self.cv_key, self.cv_map)])
        self._cscratch = None
        self._scratch = None
        self._qscratch = None
        self._bscratch = None
        self._shadow = None            # fp16 unit-normalised copy of the token rows (batched queries), built lazily
        self._shadow_valid = False
        self.next_ordinal = 0
        # small pinned rings so the per-batch parameters (poses, ordinals) and result reads never block the host
        self._ring = [None] * 8
        self._ring_ev = [None] * 8
        self._ring_i = 0
        self._cnt_ring = torch.zeros(4096, 4, dtype=torch.int32).pin_memory()
        self._cnt_i = 0
        self.reset()

    # ------------------------------------------------------------------ build
    def reset(self):
        _lib.check(self.lib.bscnav_store_reset(C.byref(self.g), C.byref(self.st), _lib.stream_ptr()), "store_reset")
        if self.col is not None:
            _lib.check(self.lib.bscnav_colour_reset(C.byref(self.g), C.byref(self.col), _lib.stream_ptr()), "colour_reset")
        self.next_ordinal = 0
        self._shadow_valid = False

    def row_norms(self, n_rows: int):
        """Recompute norm2 of the first n_rows token rows on the device (used after loading rows from disk)."""
        _lib.check(self.lib.bscnav_store_row_norms(C.byref(self.g), C.byref(self.st), int(n_rows), _lib.stream_ptr()), "row_norms")
        self._shadow_valid = False

    def _scratch_for(self, nbytes: int) -> torch.Tensor:
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._scratch

    def fuse(self, depth: torch.Tensor, poses, tokens: torch.Tensor, sample_idx: torch.Tensor | None = None,
             ord_base=None, want_records: bool = False, rgb: torch.Tensor | None = None):
        """Fuse a batch: depth (B,H,W) f32 device, poses (B,4,4) float64 (host array or device tensor),
        tokens (B, n_patch, D) f32 device; sample_idx (B, n_points) int32 device or None for dense raster order."""
        g = self.g
        B = depth.shape[0]
        self._shadow_valid = False
        assert depth.dtype == torch.float32 and depth.is_cuda and depth.is_contiguous() and depth.shape[1:] == (g.height, g.width)
        assert tokens.dtype == torch.float32 and tokens.is_contiguous() and tokens.shape == (B, g.n_patch_h * g.n_patch_w, g.dim)
        n_points = g.height * g.width if sample_idx is None else sample_idx.shape[1]
        if ord_base is None:
            if self.next_ordinal + B * n_points >= 2 ** 32:
                raise _lib.BscnavError("ordinal range exceeded (2^32 points per store)")
            ord_base = self.next_ordinal + np.arange(B, dtype=np.int64) * n_points
            self.next_ordinal += B * n_points
        # poses (B x 16 f64) and ordinals (B x u32, padded to 8 bytes) travel in one pinned staging block
        poses, ob = self._params_to_device(poses, ord_base, B)
        nbytes = _lib.check(self.lib.bscnav_fuse_scratch_bytes(B, n_points), "fuse_scratch_bytes")
        scratch = self._scratch_for(nbytes)
        rec = torch.empty(B * n_points, 4, dtype=torch.int32, device=self.device) if want_records else None
        if sample_idx is not None:
            assert sample_idx.dtype == torch.int32 and sample_idx.is_contiguous()
        _lib.check(self.lib.bscnav_fuse_frames(C.byref(g), C.byref(self.st), _lib.ptr(depth), _lib.ptr(poses), _lib.ptr(ob),
                                               _lib.ptr(sample_idx), B, n_points, _lib.ptr(tokens), _lib.ptr(scratch), nbytes,
                                               _lib.ptr(rec), _lib.stream_ptr()), "fuse_frames")
        if self.col is not None:
            # colour side channels of the same batch (same visiting order and ordinals): memory_2.py:888-903
            if rgb is None:
                raise _lib.BscnavError("this store keeps colour channels: fuse() needs the batch's rgb frames")
            assert rgb.dtype == torch.uint8 and rgb.is_cuda and rgb.is_contiguous() and rgb.shape[:3] == depth.shape
            cb = _lib.check(self.lib.bscnav_colour_scratch_bytes(C.byref(g), B, n_points), "colour_scratch_bytes")
            if self._cscratch is None or self._cscratch.numel() < cb:
                self._cscratch = torch.empty(cb, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.bscnav_colour_frames(C.byref(g), C.byref(self.st), C.byref(self.col), _lib.ptr(depth), _lib.ptr(rgb),
                                                     int(rgb.shape[3]), _lib.ptr(poses), _lib.ptr(ob), _lib.ptr(sample_idx), B,
                                                     n_points, _lib.ptr(self._cscratch), cb, _lib.stream_ptr()), "colour_frames")
        return rec

    def colour_views(self):
        """Host copies in the reference's formats: grid_rgb (max_id,3) u8, weight (max_id,) f32, cv_map (gs,gs,3) u8,
        max_height (gs,gs) f64 (-inf where no point fell)."""
        n = self.count()
        key = self.cv_key.cpu().numpy().view(np.uint64).reshape(self.g.gs, self.g.gs)
        mh = np.where(key == 0, -np.inf, ((key - np.uint64(1)) >> np.uint64(32)).astype(np.float64))
        return self.grid_rgb[:n].cpu().numpy(), self.weight[:n].cpu().numpy(), self.cv_map.cpu().numpy(), mh

    def _params_to_device(self, poses, ord_base, B):
        """One pinned staging block per batch: [poses as B*16 float64 | B uint32 ordinals] -> (poses, ordinals) on the device."""
        i = self._ring_i = (self._ring_i + 1) % len(self._ring)
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()          # only blocks if the device is a full ring behind
        buf = self._ring[i]
        if buf is None or buf.numel() < B * 33:
            buf = self._ring[i] = torch.empty(max(B, 64) * 33, dtype=torch.int32).pin_memory()
        host = buf[: B * 33].numpy()
        pose_on_device = torch.is_tensor(poses) and poses.is_cuda
        if not pose_on_device:
            host[: B * 32].view(np.float64)[:] = np.asarray(poses, dtype=np.float64).reshape(-1)
        host[B * 32:].view(np.uint32)[:] = np.asarray(ord_base, dtype=np.int64).astype(np.uint32)
        dev = buf[: B * 33].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev
        pd = poses.reshape(-1).to(torch.float64).contiguous() if pose_on_device else dev[: B * 32].view(torch.float64)
        return pd, dev[B * 32:]

    def count_async(self):
        """Enqueue a device->host read of (slots in use, overflow flag, colour mark, rows committed) into pinned memory;
        returns a handle for count_wait / rows_committed_wait."""
        i = self._cnt_i = (self._cnt_i + 1) % self._cnt_ring.shape[0]
        # written by a kernel through the mapped address of the pinned ring: no copy-engine operation on the scatter's stream
        _lib.check(self.lib.bscnav_store_publish_counts(C.byref(self.st), self._cnt_ring[i].data_ptr(), _lib.stream_ptr()),
                   "store_publish_counts")
        ev = torch.cuda.Event()
        ev.record()
        return i, ev

    def count_wait(self, handle) -> int:
        i, ev = handle
        ev.synchronize()
        c = self._cnt_ring[i].tolist()
        if c[1]:
            raise _lib.BscnavError(f"slot capacity {self.g.max_slots} exceeded (reference: IndexError at memory_2.py:891)")
        return int(c[0])

    def rows_committed_wait(self, handle) -> int:
        """Token rows copied into the store since the last reset, as of the count_async() that produced `handle`."""
        i, ev = handle
        ev.synchronize()
        return int(self._cnt_ring[i, 3]) & 0xFFFFFFFF

    def count(self) -> int:
        """Slots in use (max_id, memory_2.py:894); synchronises.  Raises if a batch overflowed the slot capacity."""
        c = self.counters[:2].tolist()
        if c[1]:
            raise _lib.BscnavError(f"slot capacity {self.g.max_slots} exceeded (reference: IndexError at memory_2.py:891)")
        return int(c[0])

    # ------------------------------------------------------------------ query
    def query(self, q: torch.Tensor, k: int = 100, region_radius=np.inf, curr_grid=None, floor_range=None):
        """q (Q, D) f32 device -> (sims (Q,k) f32, positions (Q,k,3) i32, counts (Q,) i32), device tensors."""
        g = self.g
        q = q.to(self.device, torch.float32).contiguous()
        Q = q.shape[0]
        assert q.shape[1] == g.dim
        flt = self._filter(region_radius, curr_grid, floor_range)
        nbytes = _lib.check(self.lib.bscnav_query_scratch_bytes(C.byref(g), Q, k), "query_scratch_bytes")
        if self._qscratch is None or self._qscratch.numel() < nbytes:
            self._qscratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        sims = torch.empty(Q, k, dtype=torch.float32, device=self.device)
        pos = torch.empty(Q, k, 3, dtype=torch.int32, device=self.device)
        cnt = torch.empty(Q, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.bscnav_query_topk(C.byref(g), C.byref(self.st), _lib.ptr(q), Q, k, C.byref(flt),
                                              _lib.ptr(self._qscratch), nbytes, _lib.ptr(sims), _lib.ptr(pos), _lib.ptr(cnt),
                                              _lib.stream_ptr()), "query_topk")
        return sims, pos, cnt

    def _filter(self, region_radius, curr_grid, floor_range):
        flt = _lib.Filter(radius2=-1, floor_min=1, floor_max=0, reserved=0)
        if region_radius != np.inf:
            flt.radius2 = int(np.floor(float(region_radius) ** 2))   # d2 integer: d2 <= r^2  <=>  d2 <= floor(r^2)
            flt.cur[:] = [int(v) for v in curr_grid]
        if floor_range is not None:
            flt.floor_min, flt.floor_max = int(floor_range[0]), int(floor_range[1])
            flt.reserved = 1               # explicit enable: a degenerate (empty) floor range returns nothing
        return flt

    def invalidate_shadow(self):
        """Call after writing token rows behind the store's back (tests, loaders)."""
        self._shadow_valid = False

    def query_batch(self, q: torch.Tensor, k: int = 100, region_radius=np.inf, curr_grid=None, floor_range=None, check: bool = True):
        """Many queries per pass over the store: q (Q, D) -> the same (sims, positions, counts) as `query`, bit for bit.
        Candidates come from a tcgen05 GEMM against an fp16 shadow of the rows, the result from exact fp32 re-scoring
        (csrc/query_batch.cu).  With check=True (default) the call synchronises once to read the slot count (bounds the GEMM
        to the used rows) and to re-issue, through the exact path, any query the batch path flagged (count -1)."""
        g = self.g
        q = q.to(self.device, torch.float32).contiguous()
        Q = q.shape[0]
        assert q.shape[1] == g.dim
        if self._shadow is None:
            nbytes = _lib.check(self.lib.bscnav_shadow_bytes(C.byref(g)), "shadow_bytes")
            self._shadow = torch.zeros(nbytes // 2, dtype=torch.float16, device=self.device)
        if not self._shadow_valid:
            _lib.check(self.lib.bscnav_store_build_shadow(C.byref(g), C.byref(self.st), _lib.ptr(self._shadow), _lib.stream_ptr()),
                       "build_shadow")
            self._shadow_valid = True
        hint = self.count() * g.cache if check else 0
        flt = self._filter(region_radius, curr_grid, floor_range)
        nbytes = _lib.check(self.lib.bscnav_query_batch_scratch_bytes(C.byref(g), Q, k, hint), "query_batch_scratch_bytes")
        if self._bscratch is None or self._bscratch.numel() < nbytes:
            self._bscratch = None
            self._bscratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        sims = torch.empty(Q, k, dtype=torch.float32, device=self.device)
        pos = torch.empty(Q, k, 3, dtype=torch.int32, device=self.device)
        cnt = torch.empty(Q, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.bscnav_query_topk_batch(C.byref(g), C.byref(self.st), _lib.ptr(self._shadow), _lib.ptr(q), Q, k,
                                                    C.byref(flt), hint, _lib.ptr(self._bscratch), nbytes, _lib.ptr(sims),
                                                    _lib.ptr(pos), _lib.ptr(cnt), _lib.stream_ptr()), "query_topk_batch")
        if check:
            flagged = (cnt < 0).nonzero().flatten().tolist()
            for i in flagged:                       # candidate set too large for the batch path: exact path for this query
                s1, p1, c1 = self.query(q[i:i + 1], k=k, region_radius=region_radius, curr_grid=curr_grid, floor_range=floor_range)
                sims[i], pos[i], cnt[i] = s1[0], p1[0], c1[0]
        return sims, pos, cnt

    def cluster_centers(self, sims: torch.Tensor, pos: torch.Tensor, cnt: torch.Tensor, eps: float = 10, min_samples: int = 5,
                        max_clusters: int = 32):
        """Device form of weighted_cluster_centers (BSCAgent.py:479-497) over query results (sims (Q,K), pos (Q,K,3), cnt (Q,)):
        returns (centers (Q,max_clusters,3) f64, mean_sim (Q,max_clusters) f64, sizes (Q,max_clusters) i32, labels (Q,K) i32,
        n_clusters (Q,) i32), device tensors; clusters ordered by descending mean similarity."""
        Q, K = sims.shape
        dev = self.device
        centers = torch.zeros(Q, max_clusters, 3, dtype=torch.float64, device=dev)
        mean_sim = torch.zeros(Q, max_clusters, dtype=torch.float64, device=dev)
        sizes = torch.zeros(Q, max_clusters, dtype=torch.int32, device=dev)
        labels = torch.empty(Q, K, dtype=torch.int32, device=dev)
        n_c = torch.empty(Q, dtype=torch.int32, device=dev)
        _lib.check(self.lib.bscnav_cluster_centers(_lib.ptr(pos.contiguous()), _lib.ptr(sims.contiguous()), _lib.ptr(cnt.contiguous()),
                                                   Q, K, float(eps), int(min_samples), int(max_clusters), _lib.ptr(centers),
                                                   _lib.ptr(mean_sim), _lib.ptr(sizes), _lib.ptr(labels), _lib.ptr(n_c),
                                                   _lib.stream_ptr()), "cluster_centers")
        return centers, mean_sim, sizes, labels, n_c

    # ------------------------------------------------------------------ multi-GPU deltas
    def export_delta(self, ord_lo: int, ord_hi: int, max_out: int):
        g = self.g
        dev = self.device
        d_cell = torch.full((max_out,), -1, dtype=torch.int32, device=dev)
        d_key = torch.empty(max_out, dtype=torch.int64, device=dev)
        d_h = torch.empty(max_out, dtype=torch.int32, device=dev)
        d_feat = torch.empty(max_out, g.dim, dtype=torch.float32, device=dev)
        n = torch.zeros(2, dtype=torch.int32, device=dev)     # [0] rows written, [1] rows in range (> max_out: truncated)
        nbytes = _lib.check(self.lib.bscnav_export_scratch_bytes(C.byref(g)), "export_scratch_bytes")
        scratch = self._scratch_for(nbytes)
        _lib.check(self.lib.bscnav_export_delta(C.byref(g), C.byref(self.st), ord_lo, ord_hi, _lib.ptr(d_cell), _lib.ptr(d_key),
                                                _lib.ptr(d_h), _lib.ptr(d_feat), max_out, _lib.ptr(n), _lib.ptr(scratch), nbytes,
                                                _lib.stream_ptr()), "export_delta")
        return d_cell, d_key, d_h, d_feat, n

    def export_packed(self, ord_lo: int, ord_hi: int, max_out: int, out: torch.Tensor | None = None):
        """The rows whose ordinal lies in [ord_lo, ord_hi) as ONE packed buffer (max_out, 16 + 4 D) uint8 -- the unit of the
        all-gather merge -- plus the device counter pair [rows written, rows in range]."""
        g = self.g
        rb = _lib.check(self.lib.bscnav_packed_row_bytes(C.byref(g)), "packed_row_bytes")
        if out is None:
            out = torch.empty(max_out, rb, dtype=torch.uint8, device=self.device)
        assert out.is_contiguous() and out.shape[0] >= max_out and out.shape[1] == rb
        n = torch.zeros(2, dtype=torch.int32, device=self.device)
        nbytes = _lib.check(self.lib.bscnav_export_scratch_bytes(C.byref(g)), "export_scratch_bytes")
        scratch = self._scratch_for(nbytes)
        _lib.check(self.lib.bscnav_export_packed(C.byref(g), C.byref(self.st), ord_lo, ord_hi, _lib.ptr(out), max_out, _lib.ptr(n),
                                                 _lib.ptr(scratch), nbytes, _lib.stream_ptr()), "export_packed")
        return out, n

    def merge_packed(self, packed: torch.Tensor, n_segments: int, seg_capacity: int, seg_counts: torch.Tensor, count_stride: int = 2,
                     skip_segment: int = -1):
        """Merge a gathered packed buffer (n_segments x seg_capacity rows; segment s holds seg_counts[s * count_stride] valid
        rows, read on the device) under the store's winner rule; nothing synchronises."""
        self._shadow_valid = False
        assert packed.is_contiguous() and seg_counts.dtype == torch.int32 and seg_counts.is_contiguous()
        nbytes = _lib.check(self.lib.bscnav_merge_packed_scratch_bytes(n_segments, seg_capacity), "merge_packed_scratch_bytes")
        scratch = self._scratch_for(nbytes)
        _lib.check(self.lib.bscnav_merge_packed(C.byref(self.g), C.byref(self.st), _lib.ptr(packed), n_segments, seg_capacity,
                                                _lib.ptr(seg_counts), count_stride, skip_segment, _lib.ptr(scratch), nbytes,
                                                _lib.stream_ptr()), "merge_packed")

    def merge_delta(self, d_cell, d_key, d_h, d_feat):
        self._shadow_valid = False
        n = d_cell.numel()
        chunk = 1 << 20
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            nbytes = _lib.check(self.lib.bscnav_fuse_scratch_bytes(1, e - s), "fuse_scratch_bytes")
            scratch = self._scratch_for(nbytes)
            _lib.check(self.lib.bscnav_merge_delta(C.byref(self.g), C.byref(self.st), _lib.ptr(d_cell[s:e]), _lib.ptr(d_key[s:e]),
                                                   _lib.ptr(d_h[s:e]), _lib.ptr(d_feat[s:e]), e - s, _lib.ptr(scratch), nbytes,
                                                   _lib.stream_ptr()), "merge_delta")

    # ------------------------------------------------------------------ host views (tests, I/O)
    def positions(self) -> np.ndarray:
        """(max_id, 3) int32 [row, col, h-minh] in slot order (grid_rgb_pos, memory_2.py:893).  BEV keys: h of the best token."""
        n = self.count()
        g = self.g
        cell = self.slot_cell[:n].cpu().numpy().astype(np.int64)
        if g.levels == 1:
            h = self.tok_h.view(-1, g.cache)[:n, 0].cpu().numpy()
            rc = cell
        else:
            h = cell % g.levels
            rc = cell // g.levels
        return np.stack([rc // g.gs, rc % g.gs, h], axis=1).astype(np.int32)

    def snapshot(self) -> dict:
        """Host copy of the kept tokens: {(row,col[,h]): [(r2 f32, ordinal, h, feature row)] best first}."""
        n = self.count()
        g = self.g
        C_ = g.cache
        keys = self.row_key.view(-1, C_)[:n].cpu().numpy().view(np.uint64)
        hs = self.tok_h.view(-1, C_)[:n].cpu().numpy()
        feat = self.feat.view(-1, C_, g.dim)[:n].cpu().numpy()
        pos = self.positions()
        out = {}
        for s in range(n):
            items = []
            for j in range(C_):
                k = int(keys[s, j])
                if k == _lib.KEY_EMPTY:
                    continue
                r2 = np.array([k >> 32], dtype=np.uint32).view(np.float32)[0]
                items.append((r2, k & 0xFFFFFFFF, int(hs[s, j]), feat[s, j]))
            items.sort(key=lambda t: (t[0], t[1]))
            key = (int(pos[s, 0]), int(pos[s, 1])) if g.levels == 1 else (int(pos[s, 0]), int(pos[s, 1]), int(pos[s, 2]))
            out[key] = items
        return out
