"""Small workload for compute-sanitizer (racecheck / memcheck): the smoke pair, two golden-sized cases incl. a
dominant-plane scene (DEGENSAC branch), homographies (short and long rows, elliptical correspondences), one LAF pair, one ragged batch, the matcher and the pose kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import scene_F, scene_H, scene_F_laf
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
p1, p2, _ = scene_F(500, 0.4, 3)
print("F", _cabi.fundamental_batch(p1, p2, 1.0, 0.999, iters, 0, True, 0.0, True, [7])[1].sum())
p1, p2, _ = scene_F(700, 0.3, 5, 0.8)
print("F plane", _cabi.fundamental_batch(p1, p2, 1.0, 0.999, iters, 1, True, 0.0, True, [8])[1].sum())
q1, q2, _ = scene_H(800, 300, 1)
print("H", _cabi.homography_batch(q1, q2, 3.0, 0.999, iters, 0, True, 0.0, [11])[1].sum())
q1, q2, _ = scene_H(2100, 600, 2)     # long rows: two-step residual rows (classification + dense exact list)
print("H long", _cabi.homography_batch(q1, q2, 3.0, 0.999, iters, 0, True, 0.0, [12])[1].sum())
from pydegensac_b200.scenes import scene_H2el
u10, _, _ = scene_H2el(300, 0.4, 4)
print("H 2el", _cabi.homography_2el_batch(u10, 2.0, 0.99, iters, [13])[1].sum())
l1, l2, _ = scene_F_laf(400, 0.5, 2)
print("F laf", _cabi.fundamental_batch(l1, l2, 1.0, 0.999, iters, 0, True, 3.0, True, [5])[1].sum())
print("ragged", [m.sum() for m in _cabi.fundamental_ragged([p1[:300], p1[:77], p1], [p2[:300], p2[:77], p2], 1.0, 0.99, iters, 0, True, 0.0, True, [1, 2, 3])[1]])
from pydegensac_b200.matching import match_descriptors, pose_from_fundamental
rng = np.random.default_rng(0)
d = rng.normal(size=(300, 64)).astype(np.float32)
dev = torch.device("cuda:0")
i1, i2 = match_descriptors(torch.from_numpy(d).to(dev), torch.from_numpy(d[::-1].copy() + 0.01).to(dev), ratio=0.9, mutual=True)
print("matches", len(i1))
K = torch.eye(3, dtype=torch.float64, device=dev)
F = torch.from_numpy(np.array([[0, -0.2, 0.1], [0.2, 0, -1.0], [-0.1, 1.0, 0]])).to(dev)
R, t, g = pose_from_fundamental(F[None], K, K, torch.from_numpy(p1[:100] / 500).to(dev), torch.from_numpy(p2[:100] / 500).to(dev))
print("pose", int(g[0]))
