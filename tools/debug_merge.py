import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import memory_port as mp
pkg = importlib.import_module("bsc-nav_b200")
H, W, GS, D = 96, 128, 64, 128
n, cache = 8, 4
chain = pkg.PoseChain([0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0, 0, -1, 0, 0, 0, -1], 1.5)
frames = [mp.synthetic_frame(f, H, W, n_frames=12) for f in range(n)]
depth = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
poses = np.stack([chain(f[2]) for f in frames])
tokens = torch.randn(n, 256, D, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
tok = tokens.cpu().numpy().reshape(-1, D)
cfg = pkg.StoreConfig(height=H, width=W, grid_size=GS, floor_height=-2.0, map_height=2.0, token_dim=D, cache_size=cache, bev=False)
ob = np.arange(n, dtype=np.int64) * H * W
def check(st, name):
    snap = st.snapshot(); bad = 0
    for k, items in snap.items():
        for (r2, o, h, feat) in items:
            # which token rows does each 32-float segment equal?
            segs = []
            for sgm in range(0, D, 32):
                hit = np.nonzero((tok[:, sgm:sgm+32] == feat[sgm:sgm+32]).all(axis=1))[0]
                segs.append(int(hit[0]) if len(hit) else -1)
            if len(set(segs)) != 1 or segs[0] < 0:
                bad += 1
                if bad <= 5: print(name, 'cell', k, 'key', (r2, o), 'segments from token rows', segs, 'frame of ordinal', o // (H*W))
    print(name, 'cells', len(snap), 'tokens', sum(len(v) for v in snap.values()), 'mixed rows', bad)
joint = pkg.SpatialStore(cfg); joint.fuse(depth, poses, tokens, ord_base=ob); check(joint, 'joint')
a = pkg.SpatialStore(cfg); a.fuse(depth[:3].contiguous(), poses[:3], tokens[:3].contiguous(), ord_base=ob[:3]); check(a, 'A')
b = pkg.SpatialStore(cfg); b.fuse(depth[3:].contiguous(), poses[3:], tokens[3:].contiguous(), ord_base=ob[3:]); check(b, 'B')
d_cell, d_key, d_h, d_feat, n_dev = b.export_delta(int(ob[3]), int(ob[-1]) + H*W, max_out=b.g.max_slots*cache)
k = int(n_dev.item()); print('exported', k)
df = d_feat[:k].cpu().numpy()
badx = sum(1 for i in range(k) if not (tok == df[i]).all(axis=1).any()); print('export rows not equal to any token row:', badx)
a.merge_delta(d_cell[:k].contiguous(), d_key[:k].contiguous(), d_h[:k].contiguous(), d_feat[:k].contiguous()); check(a, 'A+B')
canon = lambda s: {k: [(float(t[0]), int(t[1]), int(t[2]), t[3].tobytes()) for t in v] for k, v in s.snapshot().items()}
ca, cj = canon(a), canon(joint)
nd = 0
for k in cj:
    if ca.get(k) != cj[k]:
        nd += 1
        if nd <= 6:
            print('DIFF cell', k)
            print('  joint:', [(t[0], t[1], t[2], hash(t[3]) % 1000) for t in cj[k]])
            print('  A+B  :', [(t[0], t[1], t[2], hash(t[3]) % 1000) for t in ca.get(k, [])])
print('differing cells', nd)
