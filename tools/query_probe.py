"""Times the query paths alone at the BASELINE configs[2]/[3] shape (262 144 stored tokens, D = 1024, K = 100)."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("bsc-nav_b200")
cfg = pkg.StoreConfig(height=480, width=640, grid_size=512, token_dim=1024, cache_size=1, bev=True)
st = pkg.SpatialStore(cfg)
out = bench.query_leg(st, torch.device("cuda"), bench.peaks()[1], with_cpu=False)
print(json.dumps({k: v for k, v in out.items() if k != "workload"}, indent=1))
