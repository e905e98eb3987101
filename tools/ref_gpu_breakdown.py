"""Where does the reference's op chain spend its time on the GPU?  (one frame of cfg2; torch library kernels)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fiery_b200.synthetic import CONFIGS, make_calibration, make_head
from oracle import lift_oracle as O

cfg = CONFIGS["cfg2_static_lss"]
dev = torch.device("cuda:0")
K, E = make_calibration(cfg, seed=0)
K, E = torch.from_numpy(K).to(dev), torch.from_numpy(E).to(dev)
head = torch.from_numpy(make_head(cfg, seed=0)).to(dev)
o = O.LiftOracle.from_config(cfg).to(dev)


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


with torch.no_grad():
    ms, ego = t(lambda: o.geometry(K, E)); print(f"get_geometry            {ms:9.3f} ms")
    ms, vol = t(lambda: O.depth_context_volume(head, cfg.n_cameras, o.D, o.C)); print(f"softmax x outer product {ms:9.3f} ms")
    feats = vol[0].reshape(-1, o.C)
    ms, (idx, keep) = t(lambda: O.voxel_indices(ego[0], o.start, o.resolution, o.dimension)); print(f"voxel_indices           {ms:9.3f} ms")
    ms, f2 = t(lambda: feats[keep]); print(f"boolean index feats     {ms:9.3f} ms")
    idx2 = idx[keep]
    X, Y, Z = (int(d) for d in o.dimension)
    ranks = idx2[:, 0] * (Y * Z) + idx2[:, 1] * Z + idx2[:, 2]
    ms, order = t(lambda: ranks.argsort()); print(f"argsort                 {ms:9.3f} ms")
    ms, f3 = t(lambda: f2[order]); print(f"gather sorted feats     {ms:9.3f} ms")
    r3 = ranks[order]
    ms, cs = t(lambda: f3.cumsum(0)); print(f"cumsum(0) of (Nm,64)    {ms:9.3f} ms   Nm={f3.shape[0]}")
    ms, _ = t(lambda: O.CumsumSegmentSum.apply(f3, idx2[order], r3)); print(f"VoxelsSumming total     {ms:9.3f} ms")
    ms, _ = t(lambda: o.lift(head, K, E), n=3); print(f"whole lift, 1 frame     {ms:9.3f} ms")
