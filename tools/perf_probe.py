"""Ad-hoc timing probe (CUDA events) for the individual kernels; not the benchmark."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
vit = importlib.import_module("bsc-nav_b200.vit")

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(os.environ.get("B", 64)); T = 257; M = B * T
g = torch.Generator("cuda").manual_seed(0)
for (N, K, mode, act) in ((3072, 1024, 0, 0), (1024, 1024, 2, 0), (4096, 1024, 1, 1), (4096, 1024, 1, 0), (1024, 4096, 2, 0), (1024, 640, 4, 0)):
    a = torch.randn(M, K, device="cuda", generator=g).half(); w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16 if mode in (0, 1) else torch.float32)
    ms = timeit(lambda: vit.gemm_f16(a, w, bias, out_mode=mode, act=act, out=out))
    ref = timeit(lambda: torch.matmul(a, w.t()))
    print(f"gemm M={M} N={N} K={K} mode={mode} act={act}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s   (torch.matmul fp16: {ref:.3f} ms {2*M*N*K/ref/1e9:.1f} TF)")
cfg = pkg.CLIP_L14
eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, 0), max_batch=B)
rgb = torch.randint(0, 256, (B, 480, 640, 4), dtype=torch.uint8, device="cuda")
out = torch.empty(B, 256, 1024, device="cuda")
ms = timeit(lambda: eng.forward_frames(rgb, out=out), n=5)
print(f"vit forward B={B}: {ms:.2f} ms  {B/ms*1e3:.0f} frames/s  {cfg.flops_per_frame()*B/ms/1e9:.1f} TFLOP/s")
scfg = pkg.StoreConfig(height=480, width=640, grid_size=512, token_dim=1024, cache_size=1, bev=True)
st = pkg.SpatialStore(scfg)
from oracle import memory_port as mp
chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
frames = [mp.synthetic_frame(f, 480, 640) for f in range(B)]
depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
poses = torch.from_numpy(np.stack([chain(f[2]) for f in frames]).reshape(B, 16)).cuda()
ms = timeit(lambda: st.fuse(depth, poses, out), n=5)
print(f"fuse B={B}: {ms:.3f} ms  {B/ms*1e3:.0f} frames/s; slots {st.count()}")
q = torch.randn(8, 1024, device="cuda")
st.counters[0] = 512 * 512; st.row_key.zero_(); st.slot_cell.copy_(torch.arange(512*512, dtype=torch.int32, device="cuda")); st.occupied_ids.copy_(torch.arange(512*512, dtype=torch.int32, device="cuda"))
st.feat.normal_(); st.norm2.copy_((st.feat**2).sum(1))
for Q in (1, 8):
    ms = timeit(lambda: st.query(q[:Q], k=100), n=5)
    print(f"query Q={Q} over 262144 tokens: {ms:.3f} ms  {262144*4096/ms/1e6:.0f} GB/s algorithmic")
