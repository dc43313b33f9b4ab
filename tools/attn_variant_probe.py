"""k_attention_ws2 compile-time variants behind bscnav_debug_attention_ws2 flags: time per launch and the difference against
the product instantiation (flags 0).  FLAGS=0,32,64,96"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
lib = importlib.import_module("bsc-nav_b200._lib").lib()
B, T, H = int(os.environ.get("B", 54)), int(os.environ.get("T", 257)), 16
torch.manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * float(os.environ.get("SCALE", 0.8))).half()


def timed(flags, reps=30):
    lib.bscnav_debug_attention_ws2(flags, None)
    for _ in range(3): pkg.attention_f16(qkv, B, T, H, impl=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = pkg.attention_f16(qkv, B, T, H, impl=3)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, out.float()


base = None
for rnd in range(2):
    for f in [int(x) for x in os.environ.get("FLAGS", "0,32,64,96").split(",")]:
        us, out = timed(f)
        if base is None: base = out
        print(f"flags {f:3d}: {us:6.1f} us   max|out - product| {(out - base).abs().max().item():.3e}")
lib.bscnav_debug_attention_ws2(0, None)
