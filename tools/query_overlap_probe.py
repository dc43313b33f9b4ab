"""Would the per-query stages of one pass (select, re-score, sort, decode) overlap the candidate GEMM of the next pass?  Two stores,
two streams, 1024 queries each: aggregate rate of concurrent calls against the same calls back to back."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
cfg = pkg.StoreConfig(height=480, width=640, grid_size=512, token_dim=1024, cache_size=1, bev=True)
stores = []
for i in range(2):
    st = pkg.SpatialStore(cfg); g = st.g; n = g.max_slots
    gen = torch.Generator(device="cuda").manual_seed(7 + i)
    st.feat[:n].normal_(generator=gen)
    st.slot_cell[:n] = torch.arange(n, dtype=torch.int32, device="cuda"); st.occupied_ids[:n] = torch.arange(n, dtype=torch.int32, device="cuda")
    st.row_key[:n].zero_(); st.slot_keys[:n].zero_(); st.tok_h[:n].fill_(100); st.row_norms(n); st.counters[0] = n
    stores.append(st)
q = torch.randn(1024, 1024, device="cuda")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(concurrent, reps=6):
    for st in stores: st.query_batch(q, k=100, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for i, st in enumerate(stores):
            if concurrent:
                streams[i].wait_stream(torch.cuda.current_stream()) if r == 0 else None
                with torch.cuda.stream(streams[i]): st.query_batch(q, k=100, check=False)
            else:
                st.query_batch(q, k=100, check=False)
    if concurrent:
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{'two streams' if concurrent else 'one stream '}: {ms:.3f} ms per 2 x 1024 queries = {2048 / ms / 1e3:.2f} M queries/s", flush=True)
run(False); run(True); run(False); run(True)
