"""Block-graph debugging: run a few small forwards and report whether the engine's CUDA graph was built (stderr of the library)."""
import importlib, os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
for name, cfg in (("tiny", pkg.tiny(pkg.CLIP_L14, dim=256, depth=3, heads=4, mlp=512)), ("full", pkg.CLIP_L14)):
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=4, perturb=True), max_batch=4)
    for B in (1, 3):
        rgb = torch.randint(0, 256, (B, 96, 128, 3), dtype=torch.uint8).cuda()
        try:
            outs = []
            for rep in range(4):
                outs.append(eng.forward_frames(rgb).clone()); torch.cuda.synchronize()
            print(name, B, "ok", [torch.equal(outs[0], o) for o in outs[1:]], flush=True)
        except Exception:
            traceback.print_exc()
