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
_CPU_W = {}


def cpu_reference_rate(n_frames: int, sample_points: int, threads: int):
    """The reference's CPU algorithm on `n_frames` frames: fp32 ViT forward on `threads` host threads (its per-frame,
    batch-1 call pattern, memory_2.py:732-742) + the literal per-point loop of obs2voxeltoken (memory_2.py:863-903) over
    the first `sample_points` valid points, scaled to the frame's valid points.  Returns (frames/s, detail)."""
    from oracle import memory_port as mp
    from oracle import vit_ref
    vc = importlib.import_module("bsc-nav_b200.vit_config")
    torch.set_num_threads(threads)
    cfg = vc.CLIP_L14
    if "w" not in _CPU_W:
        _CPU_W["w"] = vc.random_weights(cfg, seed=0)
    w = _CPU_W["w"]
    p = mp.MemoryParams(height=H, width=W, grid_size=GS, token_dim=D, cache_size=1, bev=True)
    om = mp.OracleMemory(p)
    inv0 = None
    t_vit = t_loop = 0.0
    for f in range(n_frames):
        rgb, depth, pose = mp.synthetic_frame(f, H, W)
        t0 = time.perf_counter()
        x = vit_ref.preprocess_frame(torch.from_numpy(rgb[None, :, :, :3]))
        tok = vit_ref.vit_forward(x, w, cfg)[0].reshape(16, 16, D).numpy()
        t1 = time.perf_counter()
        T, inv0 = mp.pose_chain(p, pose, inv0)
        visited = mp.fuse_frame_loop(p, rgb, depth, T, tok, om, max_points=sample_points)
        t2 = time.perf_counter()
        n_valid = int(((depth > p.min_depth) & (depth < p.max_depth)).sum())
        t_vit += t1 - t0
        t_loop += (t2 - t1) * (n_valid / max(visited, 1))
    per_frame = (t_vit + t_loop) / n_frames
    detail = (f"{n_frames} frame(s) of the workload: fp32 ViT-L/14 forward on {threads} threads ({t_vit / n_frames:.2f} s/frame) + "
              f"per-point loop over the first {sample_points} of ~{H * W} points, scaled ({t_loop / n_frames:.1f} s/frame)")
    return 1.0 / per_frame, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # warm-up + K steps; each step is one frame on a bounded point sample
    vals = []
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_rate(1, args.ref_points, threads)
        if i >= args.warmup:
            vals.append(v)
        if i >= args.warmup and len(vals) >= 3 and time.time() - T_START > 240:
            break
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (ViT) / f64 (geometry)", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": detail},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
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
    for name, Q, reps in (("batch256", 256, 20), ("batch2048", 2048, 5)):
        ms = timed(lambda r: store.query_batch(q[:Q], k=100, check=False), reps)
        # shadow rows (fp16) read once per 256-query pass, similarity matrix written once and read twice
        passes = (Q + 255) // 256
        alg = passes * n * g.dim * 2 + 3 * Q * n * 4
        out[name] = {"queries_per_s": Q / (ms / 1e3), "ms_per_call": ms, "algorithmic_GBps": alg / (ms / 1e3) / 1e9,
                     "frac_of_hbm_peak": alg / (ms / 1e3) / 1e9 / peak_gbs,
                     "path": "tcgen05 candidate GEMM (fp16 shadow) + histogram select + exact fp32 re-score; device-resident queries"}

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
    p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if not os.path.exists(p):
        return None
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
    lib = blib.lib()
    B, K, Wm = args.batch, args.steps, args.warmup
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # engine + memory (the reference-facing object)
    cfg = pkg.CLIP_L14 if args.model == "clip" else pkg.DINOV2_L14_REG
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=B, device=dev)
    margs = synth.Args(height=H, width=W, grid_size=GS, depth_sample_rate=1)
    mem = pkg.VoxelTokenMemory(margs, preload_dino=eng, need_diffusion=False, cache_size=1, bev=True, max_batch=B)
    # synthetic pool: each rank owns a scene (weak scaling); P frames cycled, P*2.46 MB > L2
    P = args.pool_batches * B
    rgb_h = torch.empty(P, H, W, 4, dtype=torch.uint8, pin_memory=True)
    dep_h = torch.empty(P, H, W, dtype=torch.float32, pin_memory=True)
    poses = np.empty((P, 7))
    for f in range(P):
        r, d, po = synth.synthetic_frame(f, H, W, n_frames=1000, seed_base=100000 * rank)
        rgb_h[f] = torch.from_numpy(r)
        dep_h[f] = torch.from_numpy(d)
        poses[f] = po
    rgb_d, dep_d = rgb_h.to(dev), dep_h.to(dev)
    chain = pkg.PoseChain(margs.base_forward_axis, margs.base_left_axis, margs.base_up_axis, margs.base2cam_rot, margs.sensor_height)
    tfs = np.stack([chain(p) for p in poses]).reshape(P, 16)
    tfs_d = torch.from_numpy(tfs).to(dev)
    store = mem.store
    npts = H * W

    def step_device(s):
        # ViT of this batch on the current stream; unprojection + scatter on the memory's side stream, overlapping the
        # next batch's ViT (VoxelTokenMemory.fuse_device: the same schedule fuse_frames uses after its H2D copies)
        b0 = (s % args.pool_batches) * B
        ob = (np.arange(B, dtype=np.int64) + (s * B)) * npts % (2 ** 32 - npts * B)
        mem.fuse_device(rgb_d[b0:b0 + B], dep_d[b0:b0 + B], tfs_d[b0:b0 + B], ord_base=ob)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s in range(n):
            fn(s)
        mem._join()                    # the last batch's scatter (side stream) belongs to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    # ---- value: kernels only, inputs in HBM
    for s in range(Wm):
        step_device(s)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.bscnav_launch_count()
    ms = timed(step_device, K)
    launches = lib.bscnav_launch_count() - l0
    clocks = sampler.result()
    value = world * B * K / (ms / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM), CUDA events around every launch of two more steps
    eng.profile(True)
    for s in range(2):
        step_device(s)
    mem._join()
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    peak_tf, peak_gbs, peak_src = peaks()
    g_ms, g_flops, g_n = prof["gemm"]
    achieved = g_flops / (g_ms / 1e3) / 1e12
    step_ms_prof = sum(v[0] for v in prof.values()) / 2
    roofline = {"bound": "tensor", "kernel": "k_gemm_f16 (tcgen05, fp16 x fp16 -> fp32)", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": gemm_traffic(),
                "traffic_source": "profiles/r01_gemm_traffic.json (ncu --set full, mean over the four GEMM shapes of a layer, B=55)",
                "peak_source": peak_src,
                "launches_per_step": g_n // 2, "share_of_step": (g_ms / 2) / (ms / K),
                "breakdown_ms_per_step": {k: v[0] / 2 for k, v in prof.items()},
                "attention_tflops": prof["attention"][1] / (prof["attention"][0] / 1e3) / 1e12 if prof["attention"][0] else None}

    # ---- e2e: public API, pinned host buffers, H2D inside the timed region, D2H read of the slot count every step
    mem._clear_memory()

    reads = []

    def step_e2e(s):
        b0 = (s % args.pool_batches) * B
        mem.fuse_frames(rgb_h[b0:b0 + B], dep_h[b0:b0 + B], poses[b0:b0 + B])   # H2D on a side stream, overlapped
        reads.append(mem.max_id_async())                                         # D2H read of the step's result (pinned)

    for s in range(Wm):
        step_e2e(s)
    ms_e2e = timed(step_e2e, K)
    counts = [mem.max_id_wait(h) for h in reads[-K:]]
    assert all(c > 0 for c in counts) and counts == sorted(counts), counts
    e2e = {"value": world * B * K / (ms_e2e / 1e3), "unit": "frames/s",
           "h2d_bytes_per_step": B * (H * W * 4 + H * W * 4 + 16 * 8 + 4), "d2h_bytes_per_step": 8,
           "api": "VoxelTokenMemory.fuse_frames + max_id_async (copies overlap the previous batch)", "slots_after": counts[-1]}

    # ---- N > 1: the collective path -- treat the ranks' stores as frame shards of ONE scene: export the rows each rank produced,
    # one all-gather over NVLink, every rank merges the others' rows under the same winner rule (dist.merge_scene_across_ranks)
    merge = None
    if world > 1:
        dmod = importlib.import_module("bsc-nav_b200.dist")
        mem._join()
        slots_before = store.count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        counts = dmod.merge_scene_across_ranks(store, 0, 2 ** 32 - 1, slots_before * store.g.cache + 16)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        merge = {"ms": float(t[0]), "rows_exported_per_rank": counts, "allgather_bytes": int(sum(counts)) * (D * 4 + 16),
                 "slots_before": slots_before, "slots_after": store.count(),
                 "what": "export delta + all-gather (NCCL) + merge of the other ranks' rows; host-timed protocol steps included"}

    # every rank serves queries over its own scene (BASELINE configs[3]: scenes sharded across GPUs, no collective on the data path)
    query = query_leg(store, dev, peak_gbs, with_cpu=(world == 1 and rank == 0))
    if world > 1:
        t = torch.tensor([query["batch256"]["queries_per_s"], query["single"]["queries_per_s"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        query["all_ranks"] = {"batch256_queries_per_s": float(t[0]), "single_queries_per_s": float(t[1]), "ranks": world,
                              "sharding": "one scene per rank"}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (ViT), f64 (cell index), f32 (tokens)", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.model == "clip" else WORKLOAD.replace("CLIP ViT-L/14", "DINOv2 ViT-L/14-reg"),
                       "batch_frames_per_gpu": B, "frames_per_step": world * B,
                       "l2": f"inputs larger than L2: {B * 2 * H * W * 4 / 1e6:.0f} MB of frames per step, pool of {P} frames",
                       "sharding": "one scene per rank, no data-path collective" if world > 1 else "single GPU"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "vit_tflops": cfg.flops_per_frame() * B * K / (ms / 1e3) / 1e12, "query": query, "merge": merge}
    if world == 1:
        v, detail = cpu_reference_rate(1, args.ref_points, os.cpu_count() or 1)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": detail}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--ref-points", type=int, default=20000, help="points of each frame the CPU reference loop visits")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 54 if a.model == "clip" else 53
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
