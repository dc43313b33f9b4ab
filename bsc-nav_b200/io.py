"""Memory directory reader/writer: the reference's .npy set (memory_2.py:1136-1145, 185-200) plus the token store.

The reference keeps tokens in an HDF5 file (`feat.h5df`, group per voxel); h5py is not part of this image, so the
device store is persisted as flat arrays next to the reference's files (HDF5 interchange is a "next" row).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


def save_memory(mem, path: str):
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
    with open(path + "/long_memory.json", "w") as f:
        json.dump(mem.long_memory_dict, f, indent=4)
    rows = n * g.cache
    np.save(path + "/bscnav_slot_cell.npy", st.slot_cell[:n].cpu().numpy())
    np.save(path + "/bscnav_row_key.npy", st.row_key[:rows].cpu().numpy().view(np.uint64))
    np.save(path + "/bscnav_tok_h.npy", st.tok_h[:rows].cpu().numpy())
    np.save(path + "/bscnav_feat.npy", st.feat[:rows].cpu().numpy())
    np.save(path + "/bscnav_meta.npy", np.array([g.gs, g.levels, g.cache, g.dim, st.next_ordinal], dtype=np.int64))


def load_memory(mem, path: str):
    st = mem.store
    g = st.g
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
    mem.minh, mem.maxh = [int(v) for v in np.load(path + "/map_height.npy")]
    if os.path.exists(path + "/long_memory.json"):
        with open(path + "/long_memory.json") as f:
            mem.long_memory_dict = json.load(f)
    if hasattr(mem.Env, "original_state") and os.path.exists(path + "/original_pos.npy"):
        mem.Env.original_state.position = np.load(path + "/original_pos.npy")
    if os.path.exists(path + "/base_height.npy"):
        mem.base_height = list(np.load(path + "/base_height.npy"))
