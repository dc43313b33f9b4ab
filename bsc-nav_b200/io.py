"""Memory directory reader/writer: the reference's .npy set (memory_2.py:1136-1145, 185-200) plus the token store.

The reference keeps tokens in an HDF5 file (`feat.h5df`, one group `grid_{row}_{col}_{h}` per voxel with datasets
`features (M, D) f32` and `distances (M,) f32`, memory_2.py:335-354).  The device store is persisted natively as flat arrays
next to the reference's files; `export_h5` / `import_h5` write and read the reference's HDF5 layout -- through `h5py` when it is
importable, else through the package's own reader / writer of that file family (`h5lite.py`; h5py is not part of this image) --
so memories built by either side can be loaded by the other.  `load_single_floor` floor detection (memory_2.py:202-256) is host
logic over a handful of heights and lives here too.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


def save_memory(mem, path: str, write_h5: bool = True):
    os.makedirs(path, exist_ok=True)
    st = mem.store
    n = st.count()
    g = st.g
    np.save(path + "/grid_rgb_pos.npy", st.positions())
    np.save(path + "/max_id.npy", np.array(n))
    occ = st.occupied_ids.cpu().numpy().reshape(g.gs, g.gs, g.levels)
    np.save(path + "/occupied_ids.npy", occ)
    orig = getattr(getattr(mem.Env, "original_state", None), "position", np.zeros(3))
    np.save(path + "/original_pos.npy", np.asarray(orig))
    np.save(path + "/map_height.npy", np.array([mem.minh, mem.maxh]))
    np.save(path + "/base_height.npy", np.array(mem.base_height))
    # colour side channels of the reference's set (memory_2.py:1138-1139); zeros when the memory was built without them
    if getattr(mem, "colour", False):
        grid_rgb, weight, cv_map, _mh = st.colour_views()
        np.save(path + "/cv_map.npy", cv_map)
    else:
        grid_rgb, weight = np.zeros((n, 3), np.uint8), np.zeros(n, np.float32)
    np.save(path + "/grid_rgb.npy", grid_rgb)
    np.save(path + "/weight.npy", weight)
    with open(path + "/long_memory.json", "w") as f:
        json.dump(mem.long_memory_dict, f, indent=4)
    rows = n * g.cache
    np.save(path + "/bscnav_slot_cell.npy", st.slot_cell[:n].cpu().numpy())
    np.save(path + "/bscnav_row_key.npy", st.row_key[:rows].cpu().numpy().view(np.uint64))
    np.save(path + "/bscnav_tok_h.npy", st.tok_h[:rows].cpu().numpy())
    np.save(path + "/bscnav_feat.npy", st.feat[:rows].cpu().numpy())
    if getattr(mem, "colour", False):
        np.save(path + "/bscnav_cv_key.npy", st.cv_key.cpu().numpy().view(np.uint64))     # max_height + its ordinal (memory_2.py:901-903)
    np.save(path + "/bscnav_meta.npy", np.array([g.gs, g.levels, g.cache, g.dim, st.next_ordinal], dtype=np.int64))
    # the reference's token file last: a failure here leaves a complete native directory behind
    if g.levels > 1 and write_h5:
        export_h5(mem, path, _h5())


def _h5():
    """The HDF5 module: h5py when importable, else the package's own implementation of the same file family."""
    try:
        import h5py
        return h5py
    except ImportError:
        from . import h5lite
        return h5lite


def export_h5(mem, path: str, h5):
    """Write the kept tokens as the reference's `feat.h5df` (memory_2.py:335-354): per voxel the tokens best first."""
    st = mem.store
    g = st.g
    if g.levels == 1:
        raise ValueError("the HDF5 layout is keyed by voxel (row, col, h): not defined for a BEV-keyed store")
    # 'w': the file is rewritten as a whole (saving twice into the same directory must not fail on existing groups)
    with h5.File(os.path.join(path, "feat.h5df"), "w") as f:
        for key, items in st.snapshot().items():
            name = f"grid_{key[0]}_{key[1]}_{key[2]}"
            feats = np.stack([t[3] for t in items]).astype(np.float32)
            dist = np.array([t[0] for t in items], dtype=np.float32)
            grp = f.create_group(name)
            grp.create_dataset("features", data=feats, maxshape=(None, feats.shape[1]), chunks=True)
            grp.create_dataset("distances", data=dist, maxshape=(None,), chunks=True)


def import_h5(mem, path: str, h5):
    """Fill the device store from a reference-format directory: `feat.h5df` + `occupied_ids.npy` / `grid_rgb_pos.npy` (slot
    numbering as the reference assigned it; sequential in group order when those files are absent).  Per voxel the `cache_size`
    tokens with the smallest distances are kept (the store rule); ordinals follow group order."""
    st = mem.store
    g = st.g
    if g.levels == 1:
        raise ValueError("the HDF5 layout is keyed by voxel (row, col, h): load it into a voxel-keyed store (bev=False)")
    st.reset()
    occ = None
    if os.path.exists(path + "/occupied_ids.npy"):
        occ = np.load(path + "/occupied_ids.npy").reshape(g.gs, g.gs, g.levels)
    C_, D = g.cache, g.dim
    cells, rows_feat, rows_key, rows_h, slots = [], [], [], [], []
    ordinal = 0
    pos = np.load(path + "/grid_rgb_pos.npy") if os.path.exists(path + "/grid_rgb_pos.npy") else None
    with h5.File(os.path.join(path, "feat.h5df"), "r") as f:
        names = [nm for nm in f.keys() if nm.startswith("grid_")]
        # the reference numbers a voxel even when its token was dropped at a buffer flush (memory_2.py:880-894): slots can outnumber
        # the groups.  Size by the largest slot number in sight; slots without a group stay empty (keys KEY_EMPTY), their cell comes
        # from grid_rgb_pos.npy
        n = len(names)
        if occ is not None and occ.max() >= 0:
            n = max(n, int(occ.max()) + 1)
        if os.path.exists(path + "/max_id.npy"):
            n = max(n, int(np.load(path + "/max_id.npy")))
        feat = np.zeros((n * C_, D), np.float32)
        row_key = np.full(n * C_, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        tok_h = np.zeros(n * C_, np.int32)
        slot_cell = np.full(n, -1, np.int32)
        next_slot = int(occ.max()) + 1 if occ is not None and occ.max() >= 0 else 0
        if pos is not None:
            m = min(len(pos), n)
            slot_cell[:m] = (pos[:m, 0].astype(np.int64) * g.gs + pos[:m, 1]) * g.levels + pos[:m, 2]
        for name in names:
            row, col, h = (int(v) for v in name.split("_")[1:4])
            feats = np.asarray(f[name]["features"][:], dtype=np.float32)
            dist = np.asarray(f[name]["distances"][:], dtype=np.float32)
            if feats.shape[1] != D:
                raise ValueError(f"{name}: token dimension {feats.shape[1]} != store dimension {D}")
            slot = int(occ[row, col, h]) if occ is not None and occ[row, col, h] >= 0 else next_slot
            next_slot = max(next_slot, slot + 1)
            if slot >= n and occ is None:
                raise ValueError("slot numbering out of range")
            order = np.argsort(dist, kind="stable")[:C_]
            for j, i in enumerate(order):
                r = slot * C_ + j
                feat[r] = feats[i]
                row_key[r] = (np.uint64(np.float32(dist[i]).view(np.uint32)) << np.uint64(32)) | np.uint64(ordinal + int(i))
                tok_h[r] = h
            ordinal += len(dist)
            slot_cell[slot] = (row * g.gs + col) * g.levels + h
    n_slots = next_slot
    if n_slots > g.max_slots:
        raise ValueError(f"{n_slots} voxels exceed the store capacity {g.max_slots}")
    if (slot_cell[:n_slots] < 0).any():
        raise ValueError("a slot has neither a token group nor a position in grid_rgb_pos.npy")
    dev = st.device
    rows = n_slots * C_
    st.slot_cell[:n_slots] = torch.from_numpy(slot_cell[:n_slots]).to(dev)
    st.occupied_ids[torch.from_numpy(slot_cell[:n_slots].astype(np.int64)).to(dev)] = torch.arange(n_slots, dtype=torch.int32, device=dev)
    keys = torch.from_numpy(row_key[:rows].view(np.int64)).to(dev)
    st.row_key[:rows] = keys
    flip = keys.view(-1, C_) ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=dev)
    st.slot_keys[:rows] = (torch.sort(flip, dim=1).values ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=dev)).reshape(-1)
    st.tok_h[:rows] = torch.from_numpy(tok_h[:rows]).to(dev)
    st.feat[:rows] = torch.from_numpy(feat[:rows]).to(dev)
    st.row_norms(rows)
    st.counters[0] = n_slots
    st.next_ordinal = ordinal


def dbscan_labels(points: np.ndarray, eps: float, min_samples: int) -> np.ndarray:
    """scikit-learn's DBSCAN labelling on a small point set (host; used for the handful of base heights of load_single_floor):
    clusters = components of the core-core graph numbered by smallest core index, a border point joins the lowest-numbered
    cluster that reaches it, noise = -1 (the device kernel csrc/cluster.cu states the same rule for the top-K positions)."""
    p = np.asarray(points, dtype=np.float64).reshape(len(points), -1)
    n = len(p)
    d = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
    near = d <= eps
    core = near.sum(1) >= min_samples
    labels = np.full(n, -1, dtype=np.int64)
    cur = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        stack = [i]
        while stack:
            j = stack.pop()
            if labels[j] == -1:
                labels[j] = cur
                if core[j]:
                    stack.extend(k for k in np.nonzero(near[j])[0] if labels[k] == -1)
        cur += 1
    return labels


def detect_floor(mem, path: str | None = None):
    """memory_2.py:202-256: floors = DBSCAN(eps 0.4, min_samples n//5) clusters of the recorded base heights; the current floor's
    height range (in cells) becomes floor_min_height / floor_max_height, which voxel_localized and long_memory_filter apply."""
    base = np.asarray(mem.base_height, dtype=np.float64).reshape(-1, 1)
    if len(base) == 0:
        return
    min_samples = len(base) // 5 if len(base) // 5 > 0 else 1
    labels = dbscan_labels(base, 0.4, min_samples)
    mem.floor_heights = sorted(float(np.mean(base[labels == l])) for l in set(labels.tolist()) if l != -1)
    mem.num_floors = len(mem.floor_heights)
    if mem.num_floors == 0:
        return
    agent = getattr(mem.Env, "agent", None)
    current_height = agent.get_state().position[1] if agent is not None else mem.floor_heights[0]
    current_floor = int(np.argmin(np.abs(np.array(mem.floor_heights) - current_height)))
    pos = mem.grid_rgb_pos
    pos_range = [int(pos[:, 2].min()), int(pos[:, 2].max())]
    if mem.num_floors == 1:
        rng = pos_range
    else:
        fh, cs, ranges = mem.floor_heights, mem.cs, []
        for i in range(mem.num_floors):
            if i == 0:
                lo, hi = pos_range[0], pos_range[0] + (fh[1] - fh[0]) / cs
            elif i == mem.num_floors - 1:
                lo, hi = pos_range[0] + (fh[i] - fh[0]) / cs, pos_range[1]
            else:
                lo, hi = pos_range[0] + (fh[i] - fh[0]) / cs, pos_range[0] + (fh[i + 1] - fh[0]) / cs
            ranges.append([int(lo) + 1, int(hi) - 1])
        rng = ranges[current_floor]
    mem.floor_min_height, mem.floor_max_height = int(rng[0]), int(rng[1])
    if path is not None:
        mask = (pos[:, 2] >= mem.floor_min_height) & (pos[:, 2] <= mem.floor_max_height)
        np.save(path + f"/grid_rgb_pos_floor_{current_floor}.npy", pos[mask])
        if getattr(mem, "colour", False):
            np.save(path + f"/grid_rgb_floor_{current_floor}.npy", mem.grid_rgb[mask])


def load_memory(mem, path: str):
    st = mem.store
    g = st.g
    if not os.path.exists(path + "/bscnav_meta.npy"):
        # a directory written by the reference: tokens live in feat.h5df
        if not os.path.exists(path + "/feat.h5df"):
            raise FileNotFoundError(f"{path}: neither the native token arrays (bscnav_*.npy) nor the reference's feat.h5df")
        import_h5(mem, path, _h5())
        _load_side_files(mem, path)
        return
    meta = np.load(path + "/bscnav_meta.npy")
    if tuple(meta[:4]) != (g.gs, g.levels, g.cache, g.dim):
        raise ValueError(f"memory at {path} was built with (gs, levels, C, D) = {tuple(meta[:4])}, store is "
                         f"{(g.gs, g.levels, g.cache, g.dim)}")
    st.reset()
    cell = torch.from_numpy(np.load(path + "/bscnav_slot_cell.npy")).to(st.device)
    n = cell.numel()
    rows = n * g.cache
    st.slot_cell[:n] = cell
    st.occupied_ids[cell.long()] = torch.arange(n, dtype=torch.int32, device=st.device)
    keys = torch.from_numpy(np.load(path + "/bscnav_row_key.npy").view(np.int64)).to(st.device)
    st.row_key[:rows] = keys
    # slot_keys = the same keys ascending as unsigned 64-bit (flip the sign bit to sort as signed)
    flip = keys.view(-1, g.cache) ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=st.device)
    st.slot_keys[:rows] = (torch.sort(flip, dim=1).values ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=st.device)).reshape(-1)
    st.tok_h[:rows] = torch.from_numpy(np.load(path + "/bscnav_tok_h.npy")).to(st.device)
    feat = torch.from_numpy(np.load(path + "/bscnav_feat.npy")).to(st.device)
    st.feat[:rows] = feat
    st.row_norms(rows)
    st.counters[0] = n
    st.next_ordinal = int(meta[4])
    _load_side_files(mem, path)


def _load_colour(mem, path: str):
    """grid_rgb / weight / cv_map of a saved memory back into the device arrays (the reference restores them, memory_2.py:191-193);
    continued fusing then treats the loaded voxels as initialised (counters[2]) instead of re-initialising their colour."""
    st = mem.store
    if st.col is None or not os.path.exists(path + "/grid_rgb.npy"):
        return
    n = st.count()
    rgb, wgt = np.load(path + "/grid_rgb.npy"), np.load(path + "/weight.npy")
    m = min(n, len(rgb), len(wgt))
    st.grid_rgb[:m] = torch.from_numpy(np.ascontiguousarray(rgb[:m]).astype(np.uint8)).to(st.device)
    st.weight[:m] = torch.from_numpy(np.ascontiguousarray(wgt[:m]).astype(np.float32)).to(st.device)
    if os.path.exists(path + "/cv_map.npy"):
        st.cv_map.copy_(torch.from_numpy(np.load(path + "/cv_map.npy").astype(np.uint8)).to(st.device))
    if os.path.exists(path + "/bscnav_cv_key.npy"):
        st.cv_key.copy_(torch.from_numpy(np.load(path + "/bscnav_cv_key.npy").view(np.int64)).to(st.device))
    else:
        # a reference directory has no per-column key: rebuild max_height from the voxel positions (highest h per column); the
        # ordinal part is 0, so the next point at the same height replaces the colour, as `>=` does in the reference (:901)
        pos = st.positions().astype(np.int64)
        key = np.zeros(st.g.gs * st.g.gs, dtype=np.uint64)
        np.maximum.at(key, pos[:, 0] * st.g.gs + pos[:, 1], (pos[:, 2].astype(np.uint64) << np.uint64(32)) + np.uint64(1))
        st.cv_key.copy_(torch.from_numpy(key.view(np.int64)).to(st.device))
    st.counters[2] = m


def _load_side_files(mem, path: str):
    saved = [int(v) for v in np.load(path + "/map_height.npy")]
    if (saved[0], saved[1]) != (mem.minh, mem.maxh):
        raise ValueError(f"memory at {path} was built with height range {tuple(saved)} cells, this memory uses "
                         f"{(mem.minh, mem.maxh)} (floor_height / map_height / cell_size differ)")
    _load_colour(mem, path)
    if os.path.exists(path + "/long_memory.json"):
        with open(path + "/long_memory.json") as f:
            mem.long_memory_dict = json.load(f)
    if hasattr(mem.Env, "original_state") and os.path.exists(path + "/original_pos.npy"):
        mem.Env.original_state.position = np.load(path + "/original_pos.npy")
    if os.path.exists(path + "/base_height.npy"):
        mem.base_height = list(np.load(path + "/base_height.npy"))
    if getattr(mem.args, "load_single_floor", False):
        detect_floor(mem, path)
