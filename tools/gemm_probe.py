"""Run the four ViT-L GEMM shapes a few times (for ncu captures)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vit = importlib.import_module("bsc-nav_b200.vit")
B = int(os.environ.get("B", 64)); M = B * 257
g = torch.Generator("cuda").manual_seed(0)
shapes = ((3072, 1024, 0, 0), (1024, 1024, 2, 0), (4096, 1024, 1, 1), (1024, 4096, 2, 0))
bufs = []
for (N, K, mode, act) in shapes:
    a = torch.randn(M, K, device="cuda", generator=g).half(); w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16 if mode in (0, 1) else torch.float32)
    bufs.append((a, w, bias, out, mode, act))
for it in range(int(os.environ.get("ITERS", 3))):
    for (a, w, bias, out, mode, act) in bufs:
        vit.gemm_f16(a, w, bias, out_mode=mode, act=act, out=out)
torch.cuda.synchronize()
print("done")
