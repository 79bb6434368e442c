"""Latency of ONE config-2 pair (2000 correspondences, 30 % inliers, 10 k iterations) through the host-buffer API:
kernel time and wall time of a single findFundamentalMatrix call, against the reference on one core."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import scene_F
out = {"kernel_ms": [], "wall_ms": []}
_cabi.fundamental_batch(*scene_F(2000, 0.3, 99)[:2], 1.0, 0.9999, 10000, 0, True, 0.0, True, [99])   # warm-up
for s in range(16):
    p1, p2, _ = scene_F(2000, 0.3, s)
    t0 = time.perf_counter()
    _cabi.fundamental_batch(p1, p2, 1.0, 0.9999, 10000, 0, True, 0.0, True, [s])
    out["wall_ms"].append((time.perf_counter() - t0) * 1e3)
    out["kernel_ms"].append(_cabi.last_kernel_ms())
res = {"pairs": 16, "kernel_ms_median": float(np.median(out["kernel_ms"])), "wall_ms_median": float(np.median(out["wall_ms"])),
       "kernel_ms_min_max": [min(out["kernel_ms"]), max(out["kernel_ms"])]}
try:
    from oracle import ref
    t = []
    for s in range(4):
        p1, p2, _ = scene_F(2000, 0.3, s)
        t0 = time.perf_counter(); ref.find_fundamental(p1, p2, 1.0, 0.9999, 10000, seed=s, rng=0); t.append((time.perf_counter() - t0) * 1e3)
    res["reference_one_core_ms_median"] = float(np.median(t))
except Exception as e:
    res["reference"] = str(e)
print(json.dumps(res))
