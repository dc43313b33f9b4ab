"""Pre-processing kernel time in isolation (engine profile events), 54 frames of 640x480x4."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
cfg = pkg.tiny(pkg.CLIP_L14)
eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, 0), max_batch=54)
for ch in (4, 3):
    rgb = torch.randint(0, 256, (54, 480, 640, ch), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.forward_frames(rgb)
    eng.profile(True)
    for _ in range(10):
        eng.forward_frames(rgb)
    torch.cuda.synchronize()
    p = eng.profile_read()
    eng.profile(False)
    print("channels", ch, "preprocess ms per call", p["preprocess"][0] / 10, "launches", p["preprocess"][2])
