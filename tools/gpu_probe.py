"""First contact with the GPU: parity of the CUDA path vs the compiled reference on a few configs + a timing probe."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import scene_F, scene_H, batch_F
from oracle import ref

def nrm(M):
    n = np.linalg.norm(M)
    if n == 0: return M
    M = M / n; i = np.argmax(np.abs(M)); return M * np.sign(M.flat[i])

out = {}
t0 = time.time()
bad = 0; tot = 0
for plane in (0.0, 0.8):
    p1, p2, gt = scene_F(2000, 0.3, 0, plane)
    for seed in range(4):
        Fr, mr, sr = ref.find_fundamental(p1, p2, 1.0, 0.9999, 10000, seed=seed)
        Fg, mg, sg = _cabi.fundamental_batch(p1, p2, 1.0, 0.9999, 10000, 0, True, 0.0, True, [seed])
        md = int((mr != mg[0]).sum()); fe = float(np.linalg.norm(nrm(Fr) - nrm(Fg[0])))
        tot += 1; bad += (md != 0 or fe > 1e-6)
        print("F plane", plane, "seed", seed, sr, sg[0], "maskdiff", md, "err %.2e" % fe, flush=True)
p1, p2, gt = scene_H()
for et in range(5):
    for seed in range(2):
        Hr, mr, sr = ref.find_homography_raw(p1, p2, 3.0, 0.999, 10000, error_type=et, seed=seed)
        Hg, mg, sg = _cabi.homography_batch(p1, p2, 3.0, 0.999, 10000, et, True, 0.0, [seed])
        md = int((mr != mg[0]).sum()); fe = float(np.linalg.norm(nrm(Hr) - nrm(Hg[0])))
        tot += 1; bad += (md != 0 or fe > 1e-6)
        print("H et", et, "seed", seed, sr, sg[0], "maskdiff", md, "err %.2e" % fe, flush=True)
out["parity_bad"] = bad; out["parity_total"] = tot
print("parity bad", bad, "of", tot, "in %.1fs" % (time.time() - t0), flush=True)
# timing probe
for P in (148, 1024):
    b1, b2 = batch_F(P)
    seeds = np.arange(P, dtype=np.uint64)
    _cabi.fundamental_batch(b1[:8], b2[:8], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:8])
    t = time.time(); F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds); dt = time.time() - t
    kms = _cabi.last_kernel_ms()
    print("batch", P, "wall %.3fs kernel %.1f ms -> %.0f pairs/s (kernel)" % (dt, kms, P / (kms / 1e3)), "mean inl", m.sum(1).mean(), "LO", s[:, 1].mean(), flush=True)
    out["pairs_per_s_%d" % P] = P / (kms / 1e3)
json.dump(out, open("gpurun_out/probe.json", "w"))
