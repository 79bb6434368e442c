"""One build variant of the CUDA library: parity of the first K benchmark pairs against oracle/_ref + device rate at P pairs.

    DGLIB=tools/_prof/lib_g1.so python tools/variant.py P K [plane_frac]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import get_context
from pydegensac_b200 import _cabi
if os.environ.get("DGLIB"):
    _cabi._LIBPATH = os.path.abspath(os.environ["DGLIB"])
from pydegensac_b200.scenes import batch_F

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
plane = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0


def nrm(M):
    n = np.linalg.norm(M)
    if n == 0:
        return M
    M = M / n
    return M * np.sign(M.flat[np.argmax(np.abs(M))])


def work(args):
    lo, hi = args
    from oracle import ref
    b1, b2 = batch_F(hi - lo, 2000, 0.3, seed0=lo, plane_frac=plane)
    return [ref.find_fundamental(b1[i], b2[i], 1.0, 0.9999, 10000, seed=lo + i) for i in range(hi - lo)]


if __name__ == "__main__":
    b1, b2 = batch_F(P, 2000, 0.3, seed0=0, plane_frac=plane)
    seeds = np.arange(P, dtype=np.uint64)
    out = {"lib": os.environ.get("DGLIB"), "pairs": P, "plane_frac": plane}
    F, m, s = _cabi.fundamental_batch(b1[:max(K, 64)], b2[:max(K, 64)], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:max(K, 64)])
    if K > 0:
        nproc = min(16, os.cpu_count() or 1)
        chunks = [(i * K // nproc, (i + 1) * K // nproc) for i in range(nproc)]
        with get_context("fork").Pool(nproc) as pool:
            res = sum(pool.map(work, [c for c in chunks if c[1] > c[0]]), [])
        same = 0
        bad = []
        for i in range(K):
            Fr, mr, sr = res[i]
            ok = np.array_equal(mr, m[i]) and np.linalg.norm(nrm(Fr) - nrm(F[i])) < 1e-6 and list(sr[:2]) == list(s[i][:2])
            same += bool(ok)
            if not ok:
                bad.append(i)
        out["parity"] = "%d/%d" % (same, K)
        out["bad"] = bad[:8]
    rates = []
    for rep in range(2):
        _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
        rates.append(P / _cabi.last_kernel_ms() * 1e3)
    out["pairs_per_s"] = [round(r) for r in rates]
    print(json.dumps(out), flush=True)
