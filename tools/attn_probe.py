import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
B = int(os.environ.get("B", 55))
cfg = pkg.CLIP_L14
import dataclasses
cfg = dataclasses.replace(cfg, depth=2)
eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, 0), max_batch=B)
rgb = torch.randint(0, 256, (B, 480, 640, 4), dtype=torch.uint8, device="cuda")
for _ in range(2): eng.forward_frames(rgb)
torch.cuda.synchronize(); print("done")
