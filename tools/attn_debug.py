"""Compares the tcgen05 warp-specialised attention with the mma.sync kernel per (frame, head, query tile)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
B, T, H = int(os.environ.get("B", 3)), int(os.environ.get("T", 257)), int(os.environ.get("H", 16))
g = torch.Generator("cuda").manual_seed(T + H)
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda", generator=g) * 0.8).half()
a = pkg.attention_f16(qkv, B, T, H, impl=0).float().view(B, T, H, 64)
b = pkg.attention_f16(qkv, B, T, H, impl=2).float().view(B, T, H, 64)
err = (a - b).abs().amax(dim=-1)            # (B, T, H)
print("max err", err.max().item(), "nan", torch.isnan(b).sum().item())
for t0 in range(0, T, 32):
    e = err[:, t0:t0 + 32]
    print(f"rows {t0:3d}-{min(T, t0 + 32) - 1:3d}: max {e.max().item():9.4f}  frac>2e-3 {(e > 2e-3).float().mean().item():.3f}")
print("per head max (frame 0):", [round(v, 3) for v in err[0].amax(dim=0).tolist()])

nanrows = torch.isnan(b).any(dim=-1).nonzero()
print("nan (frame,row,head) count", len(nanrows), nanrows[:12].tolist())
bad = (err > 2e-3).nonzero()
print("bad count", len(bad), "first", bad[:8].tolist())
print("sample a", a[0, 0, 0, :4].tolist(), "b", b[0, 0, 0, :4].tolist())
print("sample a", a[0, 130, 3, :4].tolist(), "b", b[0, 130, 3, :4].tolist())
