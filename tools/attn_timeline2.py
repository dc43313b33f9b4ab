"""Where a tile's time goes in k_attention_ws2: stage-off timing experiments and the clock64 timeline of CTA 0."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
lib = importlib.import_module("bsc-nav_b200._lib").lib()
B, T, H = int(os.environ.get("B", 54)), int(os.environ.get("T", 257)), 16
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).half()


def timed(flags, reps=20):
    lib.bscnav_debug_attention_ws2(flags, None)
    for _ in range(3): pkg.attention_f16(qkv, B, T, H, impl=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): pkg.attention_f16(qkv, B, T, H, impl=3)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


names = {0: "full", 1: "no row-max pass", 2: "no exponentials", 8: "no tail rows", 16: "no epilogue", 24: "no tail, no epilogue",
         31: "nothing but loads + MMAs"}
for f, n in names.items():
    print(f"flags {f:2d} ({n}): {timed(f):.1f} us")
cols = ["a:S", "a:max", "a:P", "b:S", "b:max", "b:P", "m:Pa", "m:Oa+Sa", "m:Pb", "m:Ob+Sb", "e:start", "e:end", "t:start", "t:end", "m:Oa", "m:Ob"]
for flags in [int(x) for x in os.environ.get("TL_FLAGS", "0,31").split(",")]:
    stamps = torch.zeros(32 * 16 + 128, dtype=torch.int64, device="cuda")
    lib.bscnav_debug_attention_ws2(flags, stamps.data_ptr())
    pkg.attention_f16(qkv, B, T, H, impl=3)
    torch.cuda.synchronize()
    lib.bscnav_debug_attention_ws2(0, None)
    tt = stamps.cpu().numpy()[512:].reshape(16, 8)
    t = stamps.cpu().numpy()[:512].reshape(32, 16)
    t0 = t[t > 0].min()
    print(f"flags {flags}")
    print('tail stamps per item (start, q landed, S done, PV done, end):'); print((tt[:6, :5] - t0))
    print("tile " + " ".join(f"{c:>8s}" for c in cols))
    for n in range(int(os.environ.get("TILES", 12))):
        print(f"{n:4d} " + " ".join(f"{(t[n][i] - t0) if t[n][i] else -1:8d}" for i in range(16)))
