"""The small-M GEMMs of one frame per call (M = 257): sliced against unsliced, CUDA events over back-to-back launches inside a CUDA
graph (the stand-alone entry builds tensor maps on the host per call: eager timing would measure that)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
M = int(os.environ.get("M", 257))
scratch = torch.zeros(1024 + 152 * 128 * 128 * 4, dtype=torch.uint8, device="cuda")
for name, N, K in (("fc2", 1024, 4096), ("proj", 1024, 1024)):
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * 0.05).half()
    out = torch.zeros(M, N, device="cuda")
    for sliced in (True, False):
        fn = (lambda: pkg.gemm_f16_sliced(a, w, None, out_mode=2, out=out, scratch=scratch)) if sliced else (lambda: pkg.gemm_f16(a, w, None, out_mode=2, out=out))
        r = fn(); torch.cuda.synchronize()
        if os.environ.get("NCU"):
            continue
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn(); g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20): fn()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} M={M} N={N} K={K} {'sliced x' + str(r[1]) if sliced else 'unsliced (256-wide tiles)'}: {e0.elapsed_time(e1) / 200 * 1e3:.2f} us per launch", flush=True)
