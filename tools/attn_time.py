"""Times the two attention kernels alone (CUDA events) at the ViT-L/14 shape; not the benchmark."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
B = int(os.environ.get("B", 64)); T = int(os.environ.get("T", 257)); H = 16
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).half()
for impl in [int(x) for x in os.environ.get("IMPLS", "0,2,3").split(",")]:
    for _ in range(3): pkg.attention_f16(qkv, B, T, H, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(int(os.environ.get("REPS", 20))): pkg.attention_f16(qkv, B, T, H, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / int(os.environ.get("REPS", 20))
    print(f"attention impl={impl} B={B} T={T}: {ms*1e3:.1f} us  {4.0*B*T*T*H*64/ms/1e9:.1f} TFLOP/s")
