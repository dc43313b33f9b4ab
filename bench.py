#!/usr/bin/env python
"""Headline benchmark: RGB-D frames/s fused into a 512x512 feature grid (640x480 frames, ViT-L/14), BASELINE.json.

A step = one batch of synthetic frames through the whole build path:
    uint8 frames -> pre-process -> ViT-L/14 (tcgen05 GEMMs) -> patch tokens -> unproject + pose + cell index
    -> winner selection -> token-row commit into the grid store.
`value`  : inputs resident in HBM, kernels only (CUDA events, max over ranks).
`e2e`    : the same metric through the reference-facing API (VoxelTokenMemory.fuse_frames) with pinned HOST buffers,
           host->device copies and a device->host read of the result inside the timed region.
`--impl reference` times the reference's CPU algorithm (oracle port: torch-fp32 ViT on all host threads + the
per-point Python loop of obs2voxeltoken) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W, GS, D = 480, 640, 512, 1024
METRIC = "RGB-D frames/sec into 512² feature grid (640×480, ViT-L/14) at 1/2/4/8 B200"
WORKLOAD = ("BASELINE configs[1]: HM3D single scene, synthetic 640x480 RGB-D frames (box room, poses on a circle), "
            "CLIP ViT-L/14 seeded random weights, 512x512 BEV grid, C=1, dense sampling (rate=1)")


def _flatten(d, prefix=""):
    if isinstance(d, dict):
        for k, v in d.items():
            yield from _flatten(v, prefix + "." + str(k).lower())
    elif isinstance(d, (int, float)) and not isinstance(d, bool):
        yield prefix, float(d)


def parse_peaks(d):
    """(bf16 TFLOP/s, HBM GB/s) from a MEASURED_PEAKS.json-like dict, whatever its exact key names: the sustained bf16 figure
    is preferred over the burst one (the GEMM is timed inside a long step); values are range-checked. None where not found."""
    flat = list(_flatten(d))

    def pick(words, lo, hi, raw_scale, prefer=(), avoid=()):
        best = None
        for k, v in flat:
            if v > 1e9:                      # bytes/s or FLOP/s rather than GB/s or TFLOP/s
                v = v / raw_scale
            if not any(w in k for w in words) or not lo <= v <= hi:
                continue
            score = sum(w in k for w in prefer) - sum(w in k for w in avoid)
            if best is None or score > best[0]:
                best = (score, v)
        return None if best is None else best[1]

    tf = pick(("bf16", "tflop", "tensor", "gemm", "cublas"), 200.0, 3000.0, 1e12, prefer=("sustain", "bf16"), avoid=("burst", "fp8", "fp4"))
    gbs = pick(("hbm", "copy", "bandwidth", "dram", "gbs", "gb_s"), 2000.0, 9000.0, 1e9, prefer=("hbm", "sustain"), avoid=("burst", "l2"))
    return tf, gbs


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                tf, gbs = parse_peaks(json.load(f))
        except (OSError, ValueError):
            tf = gbs = None
        if tf is not None and gbs is not None:
            return tf, gbs, "measured (MEASURED_PEAKS.json, sustained bf16)"
        if tf is not None or gbs is not None:
            return tf or 1400.0, gbs or 6650.0, "MEASURED_PEAKS.json where present, else fallback (B200_PROFILING.md)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------- reference arm / cpu baseline
_CPU = {}


def _cpu_setup(threads: int):
    from oracle import memory_port as mp
    vc = importlib.import_module("bsc-nav_b200.vit_config")
    torch.set_num_threads(threads)
    if "w" not in _CPU:
        _CPU["w"] = vc.random_weights(vc.CLIP_L14, seed=0)
    return mp, vc.CLIP_L14, _CPU["w"]


def cpu_reference_step(frame: int, stride: int, threads: int, rate: int = 1):
    """One step of the reference's CPU algorithm on frame `frame` of the workload (oracle port; memory_2.py:842-903):
    the fp32 ViT-L/14 forward of the WHOLE frame on `threads` host threads (the reference's batch-1 call, :732-742) and the
    literal per-point Python loop of obs2voxeltoken (:863-903) over every `stride`-th point of the visiting order.
    rate > 1: the reference's own sub-sampling (depth_sample_rate, a seeded permutation; :746-749) -- then stride must be 1 and the
    whole frame is processed.  Returns a dict of measured seconds and point counts; nothing is extrapolated here."""
    from oracle import vit_ref
    mp, cfg, w = _cpu_setup(threads)
    p = mp.MemoryParams(height=H, width=W, grid_size=GS, token_dim=D, cache_size=1, bev=True, depth_sample_rate=rate)
    om = mp.OracleMemory(p)
    rgb, depth, pose = mp.synthetic_frame(frame, H, W)
    t0 = time.perf_counter()
    x = vit_ref.preprocess_frame(torch.from_numpy(rgb[None, :, :, :3]))
    tok = vit_ref.vit_forward(x, w, cfg)[0].reshape(16, 16, D).numpy()
    t1 = time.perf_counter()
    T, _ = mp.pose_chain(p, pose, None)
    perm = np.random.default_rng(frame).permutation(H * W) if rate > 1 else None
    visited = mp.fuse_frame_loop(p, rgb, depth, T, tok, om, perm=perm, stride=stride)
    t2 = time.perf_counter()
    n_visit_full = int(((depth > p.min_depth) & (depth < p.max_depth)).sum()) if rate == 1 else visited
    return {"t_vit": t1 - t0, "t_loop": t2 - t1, "visited": visited, "points_in_frame": n_visit_full}


def cpu_dense_rate(steps: int, warmup: int, stride: int, threads: int, budget_s: float):
    """frames/s of the dense (rate = 1) reference build from `steps` measured steps.  A step = whole-frame ViT + the point loop
    over 1/stride of the frame's points; the per-frame time is t_vit + t_loop * (points in frame / points visited) -- the loop is
    a flat per-point loop, its cost per point is what is measured.  Returns (value, measured wall seconds per step, steps run, text)."""
    vals, walls = [], []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m = cpu_reference_step(i, stride, threads)
        wall = time.perf_counter() - t0
        if i == 0:                                                  # the first step shows the box's speed: keep the run inside its budget
            n_all = warmup + steps
            factor = n_all * m["t_loop"] / max(budget_s - n_all * m["t_vit"], 0.25 * budget_s)
            if factor > 1.0:
                stride = int(math.ceil(stride * factor))
        if i >= warmup:
            per_frame = m["t_vit"] + m["t_loop"] * m["points_in_frame"] / max(m["visited"], 1)
            vals.append(1.0 / per_frame)
            walls.append(wall)
            last = m
        if i >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    text = (f"each step: fp32 ViT-L/14 forward of one whole frame on {threads} threads ({last['t_vit']:.2f} s) + the per-point loop over "
            f"{last['visited']} of {last['points_in_frame']} valid points (every {stride}-th, {last['t_loop']:.2f} s measured); frame time = "
            f"t_vit + t_loop x points/visited; excluded: the reference's 4 KB D2H per point (memory_2.py:883) and its HDF5 flush (:326-358)")
    return float(np.mean(vals)), float(np.mean(walls)), len(vals), text


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    stride = max(1, (H * W) // max(args.ref_points, 1))
    value, wall, n, text = cpu_dense_rate(args.steps, args.warmup, stride, threads, budget_s=args.ref_budget)
    legs = {}
    if args.ref_legs:
        # (alpha) the reference as shipped: depth_sample_rate = 1000 (args.py), whole frames, nothing scaled
        t0 = time.perf_counter()
        ms = [cpu_reference_step(100 + i, 1, threads, rate=1000) for i in range(args.ref_alpha_frames)]
        legs["alpha_as_shipped_rate1000"] = {"frames_per_s": len(ms) / (time.perf_counter() - t0), "frames": len(ms),
                                             "points_per_frame": ms[0]["visited"], "t_vit_s": float(np.mean([m["t_vit"] for m in ms])),
                                             "t_loop_s": float(np.mean([m["t_loop"] for m in ms]))}
        # (gamma) the reference's query arithmetic over stored tokens
        sample = 20000
        per = cpu_query_rate(sample)
        legs["gamma_query"] = {"queries_per_s": 1.0 / (per * GS * GS), "cores": 1,
                               "sample": f"oracle port of voxel_localized (memory_2.py:642-671) over {sample} of {GS * GS} stored tokens, per-token cost x {GS * GS}"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": n,
            "warmup": args.warmup, "ms_per_step": wall * 1e3, "frames_per_step": value * wall, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (ViT) / f64 (geometry)", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": text},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "legs": legs, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- query leg (BASELINE configs[2])
def cpu_query_rate(n_slots: int, n_queries: int = 2):
    """The reference's query arithmetic on the host (oracle port of memory_2.py:642-671: per-voxel cosine, max, sort, top-K)
    over a bounded sample of `n_slots` stored tokens; returns seconds per query per stored token."""
    from oracle import memory_port as mp
    rng = np.random.default_rng(7)
    feats = rng.standard_normal((n_slots, D), dtype=np.float32)
    keys = np.stack([np.arange(n_slots) // GS, np.arange(n_slots) % GS, np.full(n_slots, 100)], axis=1)
    slot_feats = [feats[i:i + 1] for i in range(n_slots)]
    q = rng.standard_normal((n_queries, D), dtype=np.float32)
    t0 = time.perf_counter()
    for i in range(n_queries):
        mp.query_topk(keys, slot_feats, q[i:i + 1], K=100)
    return (time.perf_counter() - t0) / n_queries / n_slots


def query_leg(store, dev, peak_gbs, with_cpu):
    """Goal queries over a fully occupied 512x512 grid (262 144 stored tokens, D=1024, K=100), BASELINE configs[2]/[3]:
    one query at a time (exact fp32 scan, HBM-bound: algorithmic bytes = N_tok*(D*4+4) per pass), and batches of 256 / 2048
    queries through the tensor-core candidate + exact re-scoring path (results bit-identical to the single-query path)."""
    g = store.g
    n = g.gs * g.gs
    gen = torch.Generator(device=dev).manual_seed(7)
    store.feat[:n].normal_(generator=gen)
    store.slot_cell[:n] = torch.arange(n, dtype=torch.int32, device=dev)
    store.occupied_ids[:n] = torch.arange(n, dtype=torch.int32, device=dev)
    store.row_key[:n].zero_()
    store.slot_keys[:n].zero_()
    store.tok_h[:n].fill_(100)
    store.row_norms(n)
    store.counters[0] = n
    q = torch.randn(2048, g.dim, device=dev, generator=torch.Generator(device=dev).manual_seed(11))
    q_host = q.cpu().pin_memory()
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            fn(r)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda r: store.query(q[r % 64:r % 64 + 1], k=100), 40)
    gbs = n * (g.dim * 4 + 4) / (ms / 1e3) / 1e9
    out["single"] = {"queries_per_s": 1e3 / ms, "ms_per_pass": ms, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak_gbs,
                     "path": "k_query_slots (exact fp32 scan) + bitonic top-K"}
    ms = timed(lambda r: store.query_batch(q[r % 64:r % 64 + 1], k=100, check=False), 40)
    out["single_via_batch_path"] = {"queries_per_s": 1e3 / ms, "ms_per_query": ms,
                                    "path": "one query through the candidate GEMM (reads the fp16 shadow: half the bytes) + exact re-score"}
    peak_tf = peaks()[0]
    for name, Q, reps in (("batch256", 256, 20), ("batch2048", 2048, 5)):
        ms = timed(lambda r: store.query_batch(q[:Q], k=100, check=False), reps)
        # algorithmic bytes: the fp16 shadow rows once per 256-query pass + the queries; algorithmic FLOPs: the candidate GEMM.
        # The pass is bound by whichever is larger at the measured peaks (the GEMM: 137 GFLOP per pass)
        passes = (Q + 255) // 256
        alg = passes * n * g.dim * 2 + Q * g.dim * 4
        flops = 2.0 * Q * n * g.dim
        t_hbm, t_tc = alg / (peak_gbs * 1e9) * 1e3, flops / (peak_tf * 1e12) * 1e3
        out[name] = {"queries_per_s": Q / (ms / 1e3), "ms_per_call": ms, "algorithmic_GBps": alg / (ms / 1e3) / 1e9,
                     "frac_of_hbm_peak": alg / (ms / 1e3) / 1e9 / peak_gbs, "gemm_tflops": flops / (ms / 1e3) / 1e12,
                     "roofline_floor_ms": max(t_hbm, t_tc), "frac_of_roofline": max(t_hbm, t_tc) / ms,
                     "path": "tcgen05 candidate GEMM over the fp16 shadow with the selection fused into its epilogue (masked fp16 "
                             "similarities + per-32-cell maxima), histogram over the maxima, exact fp32 re-score; device-resident queries"}

    def e2e_call(r):        # host queries in (pinned), host results out: the call a user of voxel_localized_batch makes
        qd = q_host[:256].to(dev, non_blocking=True)
        sims, pos, cnt = store.query_batch(qd, k=100, check=False)
        return sims.cpu(), pos.cpu(), cnt.cpu()
    ms = timed(e2e_call, 10)
    out["batch256_e2e"] = {"queries_per_s": 256 / (ms / 1e3), "ms_per_call": ms, "h2d_bytes": 256 * g.dim * 4,
                           "d2h_bytes": 256 * 100 * 16 + 256 * 4}
    if with_cpu:
        sample = 20000
        per = cpu_query_rate(sample)
        out["cpu_baseline"] = {"queries_per_s": 1.0 / (per * n), "cores": 1, "kind": "port",
                               "sample": f"oracle port of voxel_localized over {sample} of {n} stored tokens, scaled linearly"}
    out["workload"] = "BASELINE configs[2]/[3] shape: 262144 stored tokens (512x512 grid fully occupied, C=1, D=1024), K=100, cosine + top-K"
    return out


def gemm_traffic():
    """Mean DRAM traffic per k_gemm_f16 launch of a step, from the committed ncu --set full capture (profiles/)."""
    import glob
    found = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
    if not found:
        return None
    p = found[-1]                                  # the latest round's capture
    with open(p) as f:
        d = json.load(f)
    per = {l["kernel"].split()[0]: l["traffic_bytes"] for l in d["launches"]}
    try:   # one layer = qkv + proj + fc1 + fc2
        return float(per["qkv"] + per["proj"] + per["fc1"] + per["fc2"]) / 4.0
    except KeyError:
        return None


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) goes to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ge = importlib.import_module("__graft_entry__")
    blib = importlib.import_module("bsc-nav_b200._lib")
    if not os.path.exists(blib.LIB_PATH):
        if local == 0:
            ge.build()
        if world > 1:
            dist.barrier()
    pkg = importlib.import_module("bsc-nav_b200")
    synth = importlib.import_module("bsc-nav_b200.synth")
    dmod = importlib.import_module("bsc-nav_b200.dist")
    lib = blib.lib()
    B, K, Wm = args.batch, args.steps, args.warmup
    dev = torch.device("cuda", local)
    frames_mode = args.sharding == "frames"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    # the collective path is parity-checked on these very ranks before anything is timed (raises on mismatch)
    merge_check = None
    if world > 1:
        merge_check = {"ok": True, "rows_exported_per_rank": dmod.merge_self_check(),
                       "what": "frames of a small scene sharded over the ranks, packed all-gather merge == joint build, bit for bit"}

    # engine + memory (the reference-facing object)
    cfg = pkg.CLIP_L14 if args.model == "clip" else pkg.DINOV2_L14_REG
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=B, device=dev)
    margs = synth.Args(height=H, width=W, grid_size=GS, depth_sample_rate=1)
    mem = pkg.VoxelTokenMemory(margs, preload_dino=eng, need_diffusion=False, cache_size=1, bev=True, max_batch=B)
    # synthetic pool of colour / depth images (each rank its own seeds), cycled; P * 2.46 MB > L2.  The POSES never repeat
    # (synth.tour_pose of the global frame index), so every step touches new cells and commits new winner rows.
    P = args.pool_batches * B
    rgb_h = torch.empty(P, H, W, 4, dtype=torch.uint8, pin_memory=True)
    dep_h = torch.empty(P, H, W, dtype=torch.float32, pin_memory=True)
    for f in range(P):
        r, d, _ = synth.synthetic_frame(f, H, W, n_frames=1000, seed_base=100000 * (0 if frames_mode else rank))
        rgb_h[f] = torch.from_numpy(r)
        dep_h[f] = torch.from_numpy(d)
    rgb_d, dep_d = rgb_h.to(dev), dep_h.to(dev)
    store = mem.store
    npts = H * W
    n_prof = args.profile_steps

    def poses_of(f0, n):
        return np.stack([synth.tour_pose(f) for f in range(f0, f0 + n)])

    def pool_slice(f0, n):
        i = f0 % P
        return slice(i, i + n) if i + n <= P else None

    # ------------------------------------------------------------------------------------------------ scene-parallel (default)
    def make_tfs(n_frames_total):
        mem._clear_memory()
        tfs = mem._chain.batch(poses_of(0, n_frames_total)).reshape(n_frames_total, 16)
        return torch.from_numpy(tfs).to(dev)

    def step_device(s, tfs_d):
        # ViT of this batch on the current stream; unprojection + scatter on the memory's side stream, overlapping the
        # next batch's ViT (VoxelTokenMemory.fuse_device: the same schedule fuse_frames uses after its H2D copies)
        b0 = (s % args.pool_batches) * B
        f0 = s * B
        ob = (np.arange(B, dtype=np.int64) + f0) * npts % (2 ** 32 - npts * B)
        mem.fuse_device(rgb_d[b0:b0 + B], dep_d[b0:b0 + B], tfs_d[f0:f0 + B], ord_base=ob)

    def timed(fn, n, first=0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s in range(first, first + n):
            fn(s)
        mem._join()                    # the last batch's scatter (side stream) belongs to the timed region
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    peak_tf, peak_gbs, peak_src = peaks()
    line_extra = {}
    if not frames_mode:
        # ---- value: kernels only, inputs in HBM
        tfs_d = make_tfs((Wm + K + n_prof) * B)
        for s in range(Wm):
            step_device(s, tfs_d)
        mem._join()
        torch.cuda.synchronize()
        c0 = store.counters[:4].tolist()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = lib.bscnav_launch_count()
        ms = timed(lambda s: step_device(s, tfs_d), K, first=Wm)
        launches = lib.bscnav_launch_count() - l0
        clocks = sampler.result()
        c1 = store.counters[:4].tolist()
        value = world * B * K / (ms / 1e3)
        scatter_work = {"rows_committed_per_step": ((c1[3] - c0[3]) & 0xFFFFFFFF) / K, "slots_before": c0[0], "slots_after": c1[0],
                        "note": "token rows (4 KB each) copied into the grid by k_commit inside the timed region, per step per rank"}

        # ---- roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every launch of n_prof more steps (outside the
        # timed region), per-step figures, median over the steps
        eng.profile(True)
        per_step = []
        for s in range(Wm + K, Wm + K + n_prof):
            step_device(s, tfs_d)
            mem._join()
            torch.cuda.synchronize()
            per_step.append(eng.profile_read())
        eng.profile(False)
        g_tf = [p["gemm"][1] / (p["gemm"][0] / 1e3) / 1e12 for p in per_step]
        a_tf = [p["attention"][1] / (p["attention"][0] / 1e3) / 1e12 for p in per_step if p["attention"][0]]
        achieved = float(np.median(g_tf))
        brk = {k: float(np.median([p[k][0] for p in per_step])) for k in per_step[0]}
        whole = cfg.flops_per_frame() * B / (ms / K / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "k_gemm_f16 (tcgen05, fp16 x fp16 -> fp32)", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": gemm_traffic(),
                    "traffic_source": "profiles/ (ncu --set full, mean over the four GEMM shapes of a layer)",
                    "peak_source": peak_src, "profiled_steps": n_prof, "achieved_min_max": [float(min(g_tf)), float(max(g_tf))],
                    "launches_per_step": per_step[0]["gemm"][2], "share_of_step": brk["gemm"] / (ms / K),
                    "breakdown_ms_per_step": brk,
                    "attention_tflops": float(np.median(a_tf)) if a_tf else None,
                    "whole_step": {"achieved": whole, "frac": whole / peak_tf,
                                   "what": "ViT FLOPs of the step (SURVEY 8d formula) / ms_per_step of the TIMED region: every kernel of the step in the denominator"}}

        # ---- e2e: public API, pinned host buffers, H2D inside the timed region, D2H read of the counters every step
        mem._clear_memory()
        reads = []

        def step_e2e(s):
            b0 = (s % args.pool_batches) * B
            mem.fuse_frames(rgb_h[b0:b0 + B], dep_h[b0:b0 + B], poses_of(s * B, B))   # H2D on a side stream, overlapped
            reads.append(mem.max_id_async())                                          # D2H read of the step's result (pinned)

        for s in range(Wm):
            step_e2e(s)
        ms_e2e = timed(step_e2e, K, first=Wm)
        counts = [mem.max_id_wait(h) for h in reads[-K:]]
        rows = [store.rows_committed_wait(h) for h in reads[-K - 1:]]
        assert all(c > 0 for c in counts) and counts == sorted(counts), counts
        e2e = {"value": world * B * K / (ms_e2e / 1e3), "unit": "frames/s",
               "h2d_bytes_per_step": B * (H * W * 4 + H * W * 4 + 16 * 8 + 4), "d2h_bytes_per_step": 16,
               "api": "VoxelTokenMemory.fuse_frames + max_id_async (copies overlap the previous batch)",
               "slots_first_last": [counts[0], counts[-1]],
               "rows_committed_per_step": float(np.mean(np.diff(np.array(rows, dtype=np.int64)) % (2 ** 32)))}
        step_ms = ms / K
        cfg_extra = {"batch_frames_per_gpu": B, "frames_per_step": world * B,
                     "l2": f"inputs larger than L2: {B * 2 * H * W * 4 / 1e6:.0f} MB of frames per step, pool of {P} images, poses never repeat",
                     "sharding": "one scene per rank, no data-path collective" if world > 1 else "single GPU"}
        scaling = "weak"
        line_extra = {"scatter": scatter_work}
    else:
        # ------------------------------------------------------------------------------------------------ one scene, frames sharded
        # BASELINE configs[4]: the frames of ONE scene are sharded over the ranks (dist.shard_frames); each rank fuses its shard into
        # its own store, then export + ONE packed all-gather + merge (dist.merge_scene_across_ranks).  A step = the whole scene
        # (build of every shard + the merge); total work is fixed as N grows -> strong scaling.
        F = args.scene_frames
        shard = dmod.shard_frames(F, rank, world)
        all_tfs = None

        def build_scene(_s):
            nonlocal all_tfs
            mem._clear_memory()
            if all_tfs is None:
                all_tfs = torch.from_numpy(mem._chain.batch(poses_of(0, F)).reshape(F, 16)).to(dev)
            else:
                mem._chain.batch(poses_of(0, 1))              # the scene's origin is the first pose on every rank
            for f0 in range(shard.start, shard.stop, B):
                n = min(B, shard.stop - f0)
                i = f0 % (P - B + 1)
                ob = (np.arange(n, dtype=np.int64) + f0) * npts
                mem.fuse_device(rgb_d[i:i + n], dep_d[i:i + n], all_tfs[f0:f0 + n], ord_base=ob)
            mem._join()
            if world > 1:
                build_scene.counts = dmod.merge_scene_across_ranks(mem._store, shard.start * npts, shard.stop * npts, GS * GS)

        build_scene.counts = None
        for s in range(Wm):
            build_scene(s)
        sampler = ClockSampler(local)
        sampler.start()
        l0 = lib.bscnav_launch_count()
        ms = timed(build_scene, K)
        launches = lib.bscnav_launch_count() - l0
        clocks = sampler.result()
        value = F * K / (ms / 1e3)
        slots = store.count()
        # the merge alone, device-timed, on the final stores (idempotent: a second merge changes nothing)
        merge_ms = None
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            dmod.merge_scene_across_ranks(mem._store, shard.start * npts, shard.stop * npts, GS * GS)
            e1.record()
            barrier()
            merge_ms = max_over_ranks(e0.elapsed_time(e1))
        whole = cfg.flops_per_frame() * F * K / world / (ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "whole step (per rank)", "achieved": whole, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": whole / peak_tf, "traffic": None, "peak_source": peak_src}
        # ---- e2e: the same scene through the public API (dist.fuse_scene_sharded: VoxelTokenMemory.fuse_frames on pinned host
        # frames per rank, H2D inside the timed region, then the packed all-gather merge and a D2H read of the slot count)
        scene_poses = poses_of(0, F)

        def load_frames(f0, n):
            i = f0 % (P - B + 1)
            return rgb_h[i:i + n], dep_h[i:i + n]

        slots_e2e = []

        def build_scene_api(_s):
            mem._clear_memory()
            dmod.fuse_scene_sharded(mem, scene_poses, load_frames, B, max_rows=GS * GS)
            slots_e2e.append(mem.max_id_wait(mem.max_id_async()))

        build_scene_api(0)
        ms_e2e = timed(build_scene_api, K)
        assert slots_e2e[-1] == slots, (slots_e2e, slots)          # same scene, same merged store as the device-resident run
        e2e = {"value": F * K / (ms_e2e / 1e3), "unit": "frames/s",
               "h2d_bytes_per_step": F * (H * W * 4 + H * W * 4 + 16 * 8 + 4), "d2h_bytes_per_step": 16 + (8 * world if world > 1 else 0),
               "api": "dist.fuse_scene_sharded (VoxelTokenMemory.begin_shard + fuse_frames per batch, merge_scene_across_ranks) + max_id",
               "slots_after": slots_e2e[-1]}
        step_ms = ms / K
        cfg_extra = {"batch_frames_per_gpu": B, "frames_per_step": F, "scene_frames": F,
                     "l2": f"inputs larger than L2: pool of {P} images per rank, poses never repeat",
                     "sharding": f"ONE scene of {F} frames, contiguous frame shards over {world} ranks, packed NCCL all-gather merge in the timed region"}
        scaling = "strong"
        rb = 16 + 4 * D
        line_extra = {"merge": None if world == 1 else {
            "ms": merge_ms, "rows_exported_per_rank": build_scene.counts, "allgather_bytes_received_per_rank": int(max(build_scene.counts)) * world * rb,
            "slots_after": slots, "what": "export_packed + counts all-gather + ONE packed all-gather (NCCL) + merge_packed, timed alone on the final stores"}}

    # both precisions of the feature extractor: the timed run above is the fp16 throughput path; the same batch
    # through the wide-range path (split bf16 operands, three tensor-core passes, fp32 attention) for its rate and for the
    # agreement of the two token sets
    precision = None
    if not frames_mode and rank == 0 and args.wide:
        wide = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=B, device=dev, precision="bf16x3")
        tok16 = eng.forward_frames(rgb_d[:B])
        tokw = wide.forward_frames(rgb_d[:B])
        for _ in range(2):
            wide.forward_frames(rgb_d[:B], out=tokw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            wide.forward_frames(rgb_d[:B], out=tokw)
        e1.record()
        torch.cuda.synchronize()
        wms = e0.elapsed_time(e1) / 3
        precision = {"timed_path": "fp16 operands / fp32 accumulate",
                     "bf16x3": {"ms_per_batch_vit_only": wms, "vit_frames_per_s": B / (wms / 1e3),
                                "what": "wide-range parity path: 3 tensor-core passes per product, fp32 attention; 2.4e-5 per-token error vs the fp32 oracle (tests/test_gpu_precision.py)"},
                     "fp16_vs_bf16x3_max_rel_err_per_token": float(((tok16 - tokw).norm(dim=-1) / tokw.norm(dim=-1)).max()),
                     "fp16_range_overflow": bool(eng.range_overflow())}
        del wide, tokw, tok16
        torch.cuda.empty_cache()

    # every rank serves queries over its own scene (BASELINE configs[3]: scenes sharded across GPUs, no collective on the data path)
    query = None
    latency = None
    if not frames_mode:
        if args.latency and rank == 0:
            latency = latency_leg(mem, eng, synth)
        query = query_leg(store, dev, peak_gbs, with_cpu=(world == 1 and rank == 0))
        if world > 1:
            t = torch.tensor([query["batch256"]["queries_per_s"], query["single"]["queries_per_s"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            query["all_ranks"] = {"batch256_queries_per_s": float(t[0]), "single_queries_per_s": float(t[1]), "ranks": world,
                                  "sharding": "one scene per rank"}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (ViT), f64 (cell index), f32 (tokens)", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.model == "clip" else WORKLOAD.replace("CLIP ViT-L/14", "DINOv2 ViT-L/14-reg"), **cfg_extra},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "vit_tflops": roofline["whole_step"]["achieved"] if "whole_step" in roofline else roofline["achieved"],
            "query": query, "latency_b1": latency, "precision": precision, "merge_check": merge_check, **line_extra}
    if world == 1 and not frames_mode and args.cpu_baseline:
        threads = os.cpu_count() or 1
        v, wall, n, text = cpu_dense_rate(1, 0, max(1, npts // max(args.ref_points, 1)), threads, budget_s=60.0)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": text}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def latency_leg(mem, eng, synth):
    """The reference's own call pattern (memory_2.py:1086-1101): one obs2voxeltoken per simulator step, B = 1, numpy frames in.
    `sync`: host wall time per call with the device idle before and drained after (what a simulator loop that looks at the map
    every step sees); `pipelined`: back-to-back calls without a host wait (copies / ViT / scatter of consecutive frames overlap)."""
    mem._clear_memory()
    frames = [synth.synthetic_frame(f, H, W, n_frames=1000) for f in range(8)]
    n = 40
    for i in range(8):
        mem.obs2voxeltoken({"rgb": frames[i % 8][0], "depth": frames[i % 8][1]}, synth.tour_pose(i))
    mem._join()
    torch.cuda.synchronize()
    ts = []
    for i in range(8, 8 + n):
        t0 = time.perf_counter()
        mem.obs2voxeltoken({"rgb": frames[i % 8][0], "depth": frames[i % 8][1]}, synth.tour_pose(i))
        mem._join()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    for i in range(8 + n, 8 + 2 * n):
        mem.obs2voxeltoken({"rgb": frames[i % 8][0], "depth": frames[i % 8][1]}, synth.tour_pose(i))
    mem._join()
    torch.cuda.synchronize()
    pipe = (time.perf_counter() - t0) / n
    return {"api": "VoxelTokenMemory.obs2voxeltoken (B = 1, numpy obs in, H2D included)", "ms_per_frame_sync_median": float(np.median(ts) * 1e3),
            "ms_per_frame_sync_p90": float(np.percentile(ts, 90) * 1e3), "ms_per_frame_pipelined": pipe * 1e3,
            "frames_per_s_pipelined": 1.0 / pipe}


T_START = time.time()

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None,
                    help="frames per step per GPU (54 x 257 tokens = 109 row tiles = 55 CTA-pair tiles: with 74 CTA pairs every GEMM "
                         "of the layer is within 1-3 %% of a whole number of waves); default 54 for clip, 53 for dinov2 (53 x 261 tokens -> 55 pair tiles)")
    ap.add_argument("--model", default="clip", choices=["clip", "dinov2"],
                    help="clip: CLIP ViT-L/14 (BASELINE.json's configuration, default); dinov2: DINOv2 ViT-L/14 with 4 registers, the "
                         "feature extractor the reference itself loads (args.py:50)")
    ap.add_argument("--pool-batches", type=int, default=3, help="distinct batches of synthetic frames cycled through")
    ap.add_argument("--ref-points", type=int, default=10000, help="points of each frame the CPU reference loop visits per step")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds the dense leg of --impl reference may take")
    ap.add_argument("--ref-alpha-frames", type=int, default=12, help="whole frames of the as-shipped (rate = 1000) reference leg")
    ap.add_argument("--no-ref-legs", dest="ref_legs", action="store_false", help="--impl reference: skip the alpha / gamma legs")
    ap.add_argument("--profile-steps", type=int, default=10, help="steps profiled per kernel (outside the timed region) for the roofline line")
    ap.add_argument("--sharding", default="scenes", choices=["scenes", "frames"],
                    help="scenes: one scene per rank, no collective (weak scaling, the driver's run); frames: ONE scene of --scene-frames "
                         "frames sharded over the ranks, build + NCCL merge timed (strong scaling)")
    ap.add_argument("--scene-frames", type=int, default=1000)
    ap.add_argument("--no-latency", dest="latency", action="store_false", help="skip the B = 1 obs2voxeltoken latency leg")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-wide", dest="wide", action="store_false", help="skip the wide-range (bf16x3) precision leg")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 54 if a.model == "clip" else 53
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
