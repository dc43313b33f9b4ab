"""One batched query call (Q = 2048 by default, 262 144 stored tokens, K = 100): the command ncu wraps for the launch list of a pass."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
Q = int(os.environ.get("Q", 2048))
cfg = pkg.StoreConfig(height=480, width=640, grid_size=512, token_dim=1024, cache_size=1, bev=True)
st = pkg.SpatialStore(cfg)
g = st.g
n = g.max_slots
gen = torch.Generator(device="cuda").manual_seed(7)
st.feat[:n].normal_(generator=gen)
st.slot_cell[:n] = torch.arange(n, dtype=torch.int32, device="cuda")
st.occupied_ids[:n] = torch.arange(n, dtype=torch.int32, device="cuda")
st.row_key[:n].zero_(); st.slot_keys[:n].zero_(); st.tok_h[:n].fill_(100)
st.row_norms(n)
st.counters[0] = n
q = torch.randn(Q, g.dim, device="cuda", generator=gen)
for _ in range(3):
    out = st.query_batch(q, k=100)
torch.cuda.synchronize()
print("ok", out[0].shape)
