"""Share of benchmark pairs (config 2 scenes) whose CUDA result is identical to the reference's (mask equal, F within 1e-6)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import get_context
from pydegensac_b200.scenes import batch_F

P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
plane = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0

def nrm(M):
    n = np.linalg.norm(M)
    if n == 0: return M
    M = M / n; return M * np.sign(M.flat[np.argmax(np.abs(M))])

def work(args):
    lo, hi = args
    from oracle import ref
    b1, b2 = batch_F(hi - lo, 2000, 0.3, seed0=lo, plane_frac=plane)
    out = []
    for i in range(hi - lo):
        F, m, s = ref.find_fundamental(b1[i], b2[i], 1.0, 0.9999, 10000, seed=lo + i)
        out.append((F, m, s))
    return out

if __name__ == "__main__":
    from pydegensac_b200 import _cabi
    b1, b2 = batch_F(P, 2000, 0.3, seed0=0, plane_frac=plane)
    F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, np.arange(P, dtype=np.uint64))
    nproc = min(16, os.cpu_count() or 1)
    chunks = [(i * P // nproc, (i + 1) * P // nproc) for i in range(nproc)]
    with get_context("fork").Pool(nproc) as pool:
        res = sum(pool.map(work, chunks), [])
    same = 0; maskeq = 0; worst = 0.0; dI = []
    for i in range(P):
        Fr, mr, sr = res[i]
        me = np.array_equal(mr, m[i]); fe = np.linalg.norm(nrm(Fr) - nrm(F[i]))
        maskeq += me; same += (me and fe < 1e-6)
        dI.append(int(m[i].sum()) - int(mr.sum()))
    print(json.dumps({"pairs": P, "plane_frac": plane, "identical": int(same), "mask_equal": int(maskeq),
                      "mean_inlier_count_gpu_minus_ref": float(np.mean(dI)), "max_abs_inlier_count_diff": int(np.max(np.abs(dI)))}))
