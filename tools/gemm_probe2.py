import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vit = importlib.import_module("bsc-nav_b200.vit")
B = int(os.environ.get("B", 55)); M = B * 257
g = torch.Generator("cuda").manual_seed(0)
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (N, K, mode, act) in ((4096, 1024, 0, 0), (4096, 1024, 1, 1), (3072, 1024, 1, 1), (3072, 1024, 0, 0), (2048, 1024, 0, 0), (2048, 1024, 1, 1), (4096, 2048, 1, 1), (4096, 512, 1, 1)):
    a = torch.randn(M, K, device="cuda", generator=g).half(); w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ms = timeit(lambda: vit.gemm_f16(a, w, bias, out_mode=mode, act=act, out=out))
    print(f"N={N} K={K} mode={mode} act={act}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF")
