"""Small batches: the mma.sync attention kernel (row-group rounds split over CTAs) against the tcgen05 kernel (one CTA per frame and head)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
T, H = 257, 16
for B in (1, 2, 4, 6, 9, 12, 18):
    qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).half()
    row = []
    for impl in (0, 3):
        for _ in range(3): pkg.attention_f16(qkv, B, T, H, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): pkg.attention_f16(qkv, B, T, H, impl=impl)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 50 * 1e3)
    a = pkg.attention_f16(qkv, B, T, H, impl=0).float(); b = pkg.attention_f16(qkv, B, T, H, impl=3).float()
    print(f"B={B:3d}: mma.sync {row[0]:6.1f} us   tcgen05 ws2 {row[1]:6.1f} us   max diff {float((a - b).abs().max()):.2e}", flush=True)
