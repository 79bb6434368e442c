"""findHomography batch on the GPU (BASELINE config 3: 5000 correspondences, 30 % inliers, px_th 3.0, conf 0.999,
max_iters 10000): kernel rate, parity of the first pairs against the reference (Philox replay), reference CPU rate.
usage: python tools/h_rate.py [pairs=1024] [check=32]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import scene_H
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N, NIN = 5000, 1500
p1 = np.empty((P, N, 2)); p2 = np.empty((P, N, 2))
for s in range(P):
    a, b, _ = scene_H(N, NIN, s)
    p1[s], p2[s] = a, b
seeds = np.arange(P, dtype=np.uint64)
_cabi.homography_batch(p1[:32], p2[:32], 3.0, 0.999, 10000, 0, True, 0.0, seeds[:32])
H, m, st = _cabi.homography_batch(p1, p2, 3.0, 0.999, 10000, 0, True, 0.0, seeds)
ms = _cabi.last_kernel_ms()
out = {"pairs": P, "kernel_ms": ms, "pairs_per_s": P / ms * 1e3, "mean_inliers": float(m.sum(1).mean())}
try:
    from oracle import ref
    if ref.available() and C > 0:
        same = 0
        t = time.perf_counter()
        for i in range(C):
            Hr, mr, sr = ref.find_homography_raw(p1[i], p2[i], 3.0, 0.999, 10000, 0, True, 0.0, seed=int(seeds[i]), rng=ref.RNG_PHILOX)
            a = H[i] / np.linalg.norm(H[i]); b = Hr / np.linalg.norm(Hr)
            if (mr.astype(bool) == m[i].astype(bool)).all() and min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1e-6:
                same += 1
        dt = time.perf_counter() - t
        out.update({"checked": C, "identical": same, "ref_single_process_pairs_per_s": C / dt})
except Exception as ex:
    out["ref_error"] = repr(ex)
print(json.dumps(out))
