"""What the scatter on its side stream costs the step: overlapped (default), serial on one stream, ViT alone, and the ViT on a
high-priority stream with the scatter on a low-priority one."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
synth = importlib.import_module("bsc-nav_b200.synth")
H, W, GS, B, K, NB = 480, 640, 512, 54, 20, 3
dev = torch.device("cuda", 0)
cfg = pkg.CLIP_L14
eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=B, device=dev)
mem = pkg.VoxelTokenMemory(synth.Args(height=H, width=W, grid_size=GS, depth_sample_rate=1), preload_dino=eng, need_diffusion=False,
                           cache_size=1, bev=True, max_batch=B)
P = NB * B
rgb_h = torch.empty(P, H, W, 4, dtype=torch.uint8, pin_memory=True)
dep_h = torch.empty(P, H, W, dtype=torch.float32, pin_memory=True)
for f in range(P):
    r, d, _ = synth.synthetic_frame(f, H, W, n_frames=1000)
    rgb_h[f] = torch.from_numpy(r); dep_h[f] = torch.from_numpy(d)
rgb_d, dep_d = rgb_h.to(dev), dep_h.to(dev)
npts = H * W
poses_all = np.stack([synth.tour_pose(f) for f in range((K + 5) * B)])


def run(name, fn, reset=True):
    if reset: mem._clear_memory()
    for s in range(5): fn(s)
    mem._join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for s in range(5, 5 + K): fn(s)
    t_host = time.perf_counter() - t0
    mem._join(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{name:70s} {ms:7.3f} ms/step  {B / ms * 1e3:7.0f} frames/s   host enqueue {t_host / K * 1e3:6.2f} ms/step", flush=True)


mem._clear_memory()
tfs_d = torch.from_numpy(mem._chain.batch(poses_all).reshape(-1, 16)).to(dev)
def dev_step(s):
    b0 = (s % NB) * B
    mem.fuse_device(rgb_d[b0:b0 + B], dep_d[b0:b0 + B], tfs_d[s * B:(s + 1) * B], ord_base=(np.arange(B, dtype=np.int64) + s * B) * npts)
def host_step(s):
    b0 = (s % NB) * B
    mem.fuse_frames(rgb_h[b0:b0 + B], dep_h[b0:b0 + B], poses_all[s * B:(s + 1) * B])
def host_dev_inputs(s):
    b0 = (s % NB) * B
    mem.fuse_frames(rgb_d[b0:b0 + B], dep_d[b0:b0 + B], poses_all[s * B:(s + 1) * B])
reads = []
def host_step_read(s):
    host_step(s); reads.append(mem.max_id_async())
copy_stream = torch.cuda.Stream()
stage = [torch.empty(B, H, W, 4, dtype=torch.uint8, device=dev), torch.empty(B, H, W, dtype=torch.float32, device=dev)]
def dev_step_plus_copies(s):
    b0 = (s % NB) * B
    with torch.cuda.stream(copy_stream):                  # the same H2D traffic, nobody waits for it
        stage[0].copy_(rgb_h[b0:b0 + B], non_blocking=True); stage[1].copy_(dep_h[b0:b0 + B], non_blocking=True)
    dev_step(s)


def vit_only(s):
    b0 = (s % NB) * B
    eng.forward_frames(rgb_d[b0:b0 + B], out=tok)
tok = torch.empty(B, 256, cfg.dim, device=dev)
run("ViT only (no scatter)", vit_only)
run("fuse_device, overlapped (default)", dev_step)
os.environ["BSCNAV_OVERLAP"] = "0"
run("fuse_device, serial on one stream", dev_step)
os.environ["BSCNAV_OVERLAP"] = "1"
hi = torch.cuda.Stream(priority=-1)
def dev_step_hi(s):
    with torch.cuda.stream(hi):
        dev_step(s)
def run_hi(name, fn):
    with torch.cuda.stream(hi):
        run(name, fn)
run_hi("fuse_device, ViT stream high priority, scatter stream default", dev_step)
run("fuse_device, overlapped (default, again)", dev_step)
