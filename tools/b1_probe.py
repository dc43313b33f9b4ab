"""One frame per call: where the time of a ViT-L/14 forward at batch 1 goes (per kernel class, CUDA events around every launch)
and the host-side cost of enqueueing it."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
B = int(os.environ.get("B", 1))
cfg = pkg.CLIP_L14
eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=0), max_batch=max(B, 1))
rgb = torch.randint(0, 255, (B, 480, 640, 4), dtype=torch.uint8, device="cuda")
out = eng.forward_frames(rgb)
for _ in range(5): eng.forward_frames(rgb, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
N = 50
for _ in range(N): eng.forward_frames(rgb, out=out)
t_host = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / N:.3f} ms per forward back to back (device), host enqueue {t_host / N * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(N):
    eng.forward_frames(rgb, out=out); torch.cuda.synchronize()
print(f"B={B}: {(time.perf_counter() - t0) / N * 1e3:.3f} ms per forward, synchronised after each")
eng.profile(True)
for _ in range(3):
    eng.forward_frames(rgb, out=out); torch.cuda.synchronize(); p = eng.profile_read()
eng.profile(False)
print({k: (round(v[0], 4), v[2]) for k, v in p.items()}, "ms, launches; sum", round(sum(v[0] for v in p.values()), 3))
