"""GPU: the schedule bench.py times -- VoxelTokenMemory.fuse_frames over several batches (H2D on a copy stream, ViT on the
main stream, unprojection + scatter on a side stream, double-buffered inputs and tokens) -- at the BASELINE configs[1] shape
(640x480 frames, full CLIP ViT-L/14 engine, 512x512 BEV grid, C = 1, dense), against

  * the numpy oracle (oracle/memory_port.py) fed the GPU engine's own token rows: cell set, slot order, winner keys, ordinals,
    heights and stored rows bit-exact (memory_2.py:842-903 under the store rule of SURVEY 8a);
  * a serial single-stream build from device tensors (BSCNAV_OVERLAP=0): every store array bit-identical;
  * oracle/vit_ref.py (fp32 torch) for the tokens themselves: 1e-3 relative per token;
  * the port's top-K query over the oracle's store.
"""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import memory_port as mp
from oracle import vit_ref

pytestmark = pytest.mark.gpu

H, W, GS, NB, B = 480, 640, 512, 4, 8


@pytest.fixture(scope="module")
def scene(pkg):
    synth = importlib.import_module("bsc-nav_b200.synth")
    cfg = pkg.CLIP_L14
    w = pkg.random_weights(cfg, seed=0)
    eng = pkg.ViTEngine(cfg, w, max_batch=B)
    frames = [synth.synthetic_frame(f, H, W, n_frames=NB * B) for f in range(NB * B)]   # a full turn: every batch sees new cells
    rgb = np.stack([f[0] for f in frames])
    depth = np.stack([f[1] for f in frames])
    poses = np.stack([f[2] for f in frames])
    args = synth.Args(height=H, width=W, grid_size=GS, depth_sample_rate=1)
    return dict(cfg=cfg, w=w, eng=eng, rgb=rgb, depth=depth, poses=poses, args=args)


def _arrays(mem):
    st = mem.store
    n = st.count()
    torch.cuda.synchronize()
    return n, dict(occupied_ids=st.occupied_ids.clone(), slot_cell=st.slot_cell[:n].clone(), slot_keys=st.slot_keys[:n].clone(),
                   row_key=st.row_key[:n].clone(), tok_h=st.tok_h[:n].clone(), norm2=st.norm2[:n].clone(), feat=st.feat[:n].clone())


@pytest.fixture(scope="module")
def overlapped(pkg, scene):
    """The benched schedule, from numpy inputs in a tight loop, with the device several batches behind the host."""
    mem = pkg.VoxelTokenMemory(scene["args"], preload_dino=scene["eng"], need_diffusion=False, cache_size=1, bev=True, max_batch=B)
    a = torch.randn(8192, 8192, device="cuda")
    for _ in range(30):                                # ~100 ms of queued work: the host runs ahead through all four batches
        a = (a @ a).clamp_(-1, 1)
    for b in range(NB):
        s = slice(b * B, (b + 1) * B)
        mem.fuse_frames(scene["rgb"][s], scene["depth"][s], scene["poses"][s])
    n, arr = _arrays(mem)
    return mem, n, arr


def test_overlapped_equals_serial_build(pkg, scene, overlapped):
    _, n, arr = overlapped
    os.environ["BSCNAV_OVERLAP"] = "0"
    try:
        mem = pkg.VoxelTokenMemory(scene["args"], preload_dino=scene["eng"], need_diffusion=False, cache_size=1, bev=True, max_batch=B)
        for b in range(NB):
            s = slice(b * B, (b + 1) * B)
            mem.fuse_frames(torch.from_numpy(scene["rgb"][s]).cuda(), torch.from_numpy(scene["depth"][s]).cuda(), scene["poses"][s])
            torch.cuda.synchronize()
        n2, arr2 = _arrays(mem)
    finally:
        del os.environ["BSCNAV_OVERLAP"]
    assert n == n2 and n > 3000
    for k in arr:
        assert torch.equal(arr[k], arr2[k]), k


def test_overlapped_build_matches_oracle(scene, overlapped):
    mem, n, arr = overlapped
    eng = scene["eng"]
    # the engine's own tokens, same batch composition as the build (bit-identical by construction: deterministic kernels)
    tok = torch.cat([eng.forward_frames(torch.from_numpy(scene["rgb"][b * B:(b + 1) * B]).cuda()).cpu() for b in range(NB)]).numpy()
    p = mp.MemoryParams(height=H, width=W, grid_size=GS, token_dim=scene["cfg"].dim, cache_size=1, bev=True)
    inv0, frames = None, []
    first = {}
    for f in range(NB * B):
        T, inv0 = mp.pose_chain(p, scene["poses"][f], inv0)
        rec = mp.fuse_records(p, scene["depth"][f], T)
        frames.append((f, rec))
        cell = rec["row"] * GS + rec["col"]
        u, idx = np.unique(cell, return_index=True)                  # first visit of each cell in this frame (raster order)
        for c, i in zip(u.tolist(), idx.tolist()):
            if c not in first:
                first[c] = f * H * W + int(rec["pix"][i])
    want = mp.kept_vectorised(p, frames)
    assert n == len(want)
    # slot numbering == first-touch order (memory_2.py:888-894)
    order = sorted(first, key=first.get)
    assert np.array_equal(arr["slot_cell"].cpu().numpy(), np.array(order, dtype=np.int32))
    keys = arr["row_key"].cpu().numpy().view(np.uint64)
    hs = arr["tok_h"].cpu().numpy()
    feat = arr["feat"].cpu().numpy()
    occ = arr["occupied_ids"].cpu().numpy()
    for (row, col), items in want.items():
        r2f, ordinal, (fid, ppy, ppx), h = items[0]
        s = occ[row * GS + col]
        assert s >= 0
        k = int(keys[s])
        assert np.array([k >> 32], dtype=np.uint32).view(np.float32)[0] == r2f and (k & 0xFFFFFFFF) == ordinal, (row, col)
        assert hs[s] == h
        assert np.array_equal(feat[s], tok[fid, ppy * 16 + ppx]), (row, col)
    # tokens themselves vs the fp32 torch oracle (two frames of different batches)
    for f in (3, 2 * B + 5):
        ref = vit_ref.vit_forward(vit_ref.preprocess_frame(torch.from_numpy(scene["rgb"][f:f + 1, :, :, :3])), scene["w"], scene["cfg"])[0].numpy()
        rel = np.linalg.norm(tok[f] - ref, axis=-1) / np.linalg.norm(ref, axis=-1)
        assert rel.max() < 1e-3, rel.max()
    # top-K query over the built store vs the port
    cells = sorted(want)
    keys3 = np.array([(r, c, want[(r, c)][0][3]) for r, c in cells])
    slot_feats = [tok[want[k][0][2][0], want[k][0][2][1] * 16 + want[k][0][2][2]][None] for k in cells]
    q = tok[5, 77] + 0.05 * np.random.default_rng(3).standard_normal(scene["cfg"].dim).astype(np.float32)
    _, opos, osim = mp.query_topk(keys3, slot_feats, q[None], K=50)
    sims, pos, cnt = mem.store.query(torch.from_numpy(q[None]).cuda(), k=50)
    assert int(cnt[0]) == 50
    assert np.array_equal(pos[0].cpu().numpy().astype(np.int64), opos)
    np.testing.assert_allclose(sims[0].cpu().numpy(), osim, atol=2e-6)
    sims_b, pos_b, cnt_b = mem.store.query_batch(torch.from_numpy(q[None]).cuda(), k=50)
    assert torch.equal(pos_b, pos) and torch.equal(sims_b, sims)


def test_numpy_tight_loop_equals_device_path(pkg):
    """ADVICE r1 (pinned staging race): many small batches from numpy arrays, issued without any synchronisation while the
    device is busy, must give the store the device-tensor path gives."""
    synth = importlib.import_module("bsc-nav_b200.synth")
    h, w, gs, nb, bsz = 96, 128, 64, 24, 3
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=2, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=2, perturb=True), max_batch=bsz)
    args = synth.Args(height=h, width=w, grid_size=gs, depth_sample_rate=1, floor_height=-2.0, map_height=2.0)
    frames = [synth.synthetic_frame(f, h, w, n_frames=nb * bsz) for f in range(nb * bsz)]
    rgb, depth, poses = (np.stack([f[i] for f in frames]) for i in range(3))
    out = []
    for mode in ("numpy", "device"):
        mem = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=1, max_batch=bsz, max_slots=65536)
        if mode == "numpy":
            a = torch.randn(8192, 8192, device="cuda")
            for _ in range(30):
                a = (a @ a).clamp_(-1, 1)
        for b in range(nb):
            s = slice(b * bsz, (b + 1) * bsz)
            if b == nb - 1:
                s = slice(b * bsz, b * bsz + 2)                      # ragged last batch: buffers are views, nothing is re-allocated
            if mode == "numpy":
                mem.fuse_frames(rgb[s].copy(), depth[s].copy(), poses[s])     # temporaries: only the staging copy keeps the data
            else:
                mem.fuse_frames(torch.from_numpy(rgb[s]).cuda(), torch.from_numpy(depth[s]).cuda(), poses[s])
                torch.cuda.synchronize()
        out.append(_arrays(mem))
    (n1, a1), (n2, a2) = out
    assert n1 == n2 and n1 > 100
    for k in a1:
        assert torch.equal(a1[k], a2[k]), k


def _canon_store(st):
    """Order-independent view of a store: {cell: [(r^2, ordinal, height, checksum of the token row)]} (slot numbers differ
    between a joint build and a merged one; the kept-token set may not)."""
    return {k: [(float(t[0]), int(t[1]), int(t[2]), float(np.dot(t[3].astype(np.float64), np.arange(1, t[3].size + 1)))) for t in v]
            for k, v in st.snapshot().items()}


@pytest.mark.parametrize("cache,bev", [(1, True), (3, False)])
def test_sharded_build_through_the_facade_equals_joint_build(pkg, cache, bev):
    """BASELINE configs[4] through the public API: two memories each fuse a contiguous shard of one scene's frames
    (VoxelTokenMemory.begin_shard + fuse_frames, the calls dist.fuse_scene_sharded makes on every rank), exchange their packed
    rows, and both end with exactly the joint build's kept tokens.  The second shard's pose chain must hang on the scene's first
    pose and its points must be numbered after the first shard's (ties between shards go to the earlier frame)."""
    synth = importlib.import_module("bsc-nav_b200.synth")
    h, w, gs, n, bsz = 96, 128, 64, 12, 3
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=2, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=2, perturb=True), max_batch=bsz)
    args = synth.Args(height=h, width=w, grid_size=gs, depth_sample_rate=1, floor_height=-2.0, map_height=2.0)
    frames = [synth.synthetic_frame(f % 5, h, w, n_frames=9) for f in range(n)]     # frames 5.. repeat 0..: exact r^2 ties across shards
    rgb, depth, poses = (np.stack([f[i] for f in frames]) for i in range(3))

    def new():
        return pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=cache, bev=bev, max_batch=bsz, max_slots=65536)

    joint = new()
    for f0 in range(0, n, bsz):
        joint.fuse_frames(rgb[f0:f0 + bsz], depth[f0:f0 + bsz], poses[f0:f0 + bsz])
    joint._join()
    cut = 6
    shards = []
    for lo, hi in ((0, cut), (cut, n)):
        m = new()
        m.begin_shard(poses[0], lo)
        for f0 in range(lo, hi, bsz):
            m.fuse_frames(rgb[f0:f0 + bsz], depth[f0:f0 + bsz], poses[f0:f0 + bsz])
        m._join()
        shards.append((m, lo * h * w, hi * h * w))
    packed = [m.store.export_packed(lo, hi, 65536) for m, lo, hi in shards]
    cap = max(int(nd[0]) for _, nd in packed)
    rows = torch.cat([p[:cap] for p, _ in packed])
    counts = torch.cat([nd for _, nd in packed])
    ref = _canon_store(joint.store)
    assert len(ref) > 100
    for r, (m, _, _) in enumerate(shards):
        m.store.merge_packed(rows, 2, cap, counts, count_stride=2, skip_segment=r)
        got = _canon_store(m.store)
        bad = [k for k in set(ref) | set(got) if ref.get(k) != got.get(k)]
        assert not bad, (r, len(bad), bad[:3], [ref.get(k) for k in bad[:2]], [got.get(k) for k in bad[:2]])
        # the asynchronous count issued after the merge sees the merged store (it must not ride the scatter stream past the merge)
        assert m.max_id_wait(m.max_id_async()) == joint.max_id == m.max_id


def test_max_id_async_per_batch_matches_synchronous_counts(pkg):
    """bench.py's e2e leg reads max_id after every batch without stalling the pipeline (the copy rides the scatter's stream):
    every handle must report the slot count a synchronous build reports after the same batch, and a reset issued right after
    an asynchronous read must not disturb it."""
    synth = importlib.import_module("bsc-nav_b200.synth")
    h, w, gs, nb, bsz = 96, 128, 64, 10, 3
    cfg = pkg.tiny(pkg.CLIP_L14, dim=128, depth=2, heads=2, mlp=256)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=2, perturb=True), max_batch=bsz)
    args = synth.Args(height=h, width=w, grid_size=gs, depth_sample_rate=1, floor_height=-2.0, map_height=2.0)
    frames = [synth.synthetic_frame(f, h, w, n_frames=nb * bsz) for f in range(nb * bsz)]
    rgb, depth, poses = (np.stack([f[i] for f in frames]) for i in range(3))
    sync = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=1, max_batch=bsz, max_slots=65536)
    want = []
    for b in range(nb):
        s = slice(b * bsz, (b + 1) * bsz)
        sync.fuse_frames(rgb[s], depth[s], poses[s])
        want.append(sync.max_id)
    mem = pkg.VoxelTokenMemory(args, preload_dino=eng, need_diffusion=False, cache_size=1, max_batch=bsz, max_slots=65536)
    handles = []
    for b in range(nb):
        s = slice(b * bsz, (b + 1) * bsz)
        mem.fuse_frames(rgb[s], depth[s], poses[s])
        handles.append(mem.max_id_async())
    mem._clear_memory()                                     # queued behind the last asynchronous read
    assert [mem.max_id_wait(hd) for hd in handles] == want and want == sorted(want) and want[-1] > want[0] > 0
    assert mem.max_id == 0
