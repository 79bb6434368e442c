"""Phase counters of the -DDG_PROF build on the H path (config 3)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
_cabi._LIBPATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), sys.argv[1] if len(sys.argv) > 1 else "tools/_prof/libdegensac_b200_prof.so")
from pydegensac_b200.scenes import scene_H
P = int(sys.argv[2]) if len(sys.argv) > 2 else 592
N = 5000
p1 = np.empty((P, N, 2)); p2 = np.empty((P, N, 2))
for s in range(P):
    a, b, _ = scene_H(N, 1500, s)
    p1[s], p2[s] = a, b
seeds = np.arange(P, dtype=np.uint64)
L = _cabi.lib()
_cabi.homography_batch(p1[:8], p2[:8], 3.0, 0.999, 10000, 0, True, 0.0, seeds[:8])
buf = (ctypes.c_ulonglong * 64)()
L.dgb200_prof_read(buf, 1)
H, m, st = _cabi.homography_batch(p1, p2, 3.0, 0.999, 10000, 0, True, 0.0, seeds)
ms = _cabi.last_kernel_ms()
L.dgb200_prof_read(buf, 1)
print("pairs", P, "kernel ms %.1f -> %.0f pairs/s" % (ms, P / ms * 1e3), "samples drawn/pair %.0f  LO runs/pair %.2f" % (st[:, 0].mean(), st[:, 1].mean()))
per = lambda i, n: buf[i] / max(buf[n], 1)
print("hash %.2f Mcyc/pair" % (buf[4] / P / 1e6))
print("inlidxs/pair %.0f (%.1f kcyc each, %.2f Mcyc/pair)" % (buf[22] / P, per(21, 22) / 1e3, buf[21] / P / 1e6))
print("randsubset %.2f Mcyc/pair" % (buf[23] / P / 1e6))
print("eig solver %.2f Mcyc/pair" % (buf[29] / P / 1e6))
print("total CTA cycles/pair at 2 CTAs/SM ~ %.1f M" % (296 / (P / ms * 1e3) * 1.965e9 / 1e6))
