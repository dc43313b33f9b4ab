"""Per-tile timeline of CTA 0 of the warp-specialised attention kernel (clock64 stamps), to see where a tile's time goes."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
lib = importlib.import_module("bsc-nav_b200._lib").lib()
B, T, H = 54, 257, 16
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).half()
for _ in range(3): pkg.attention_f16(qkv, B, T, H, impl=2)
torch.cuda.synchronize()
if not hasattr(lib, "bscnav_debug_attention_timeline"):
    sys.exit("this needs the instrumented build of csrc/attention_ws.cu (WS_STAMP clock64 stamps + bscnav_debug_attention_timeline): "
             "see DESIGN.md section 3.1; the shipped kernel carries no instrumentation")
lib.bscnav_debug_attention_timeline.argtypes = [C.c_int, C.c_void_p]
lib.bscnav_debug_attention_timeline(1, None)
pkg.attention_f16(qkv, B, T, H, impl=2)
torch.cuda.synchronize()
buf = np.zeros(16 * 64, dtype=np.int64)
lib.bscnav_debug_attention_timeline(0, buf.ctypes.data)
t = buf.reshape(64, 16)
t0 = t[0][t[0] > 0].min()
names = ["a:S ready", "a:pass1+max", "a:pass2 end", "a:arrived", "b:S ready", "b:pass1+max", "b:pass2 end", "b:arrived",
         "mma:p_a seen", "mma:o_empty ok", "mma:Oa+Sa issued", "mma:p_b seen", "mma:Ob+Sb issued", "epi:start", "epi:end", "mma:Oa issued"]
print("tile " + " ".join(f"{n:>16s}" for n in names))
for n in range(int(os.environ.get('TILES', 7))):
    print(f"{n:4d} " + " ".join(f"{(t[n][i] - t0) if t[n][i] else -1:16d}" for i in range(16)))
