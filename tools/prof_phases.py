"""Phase breakdown with the -DDG_PROF build (gpurun_out/libdegensac_b200_prof.so): thread-0 clock64 sums per phase."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
_cabi._LIBPATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/libdegensac_b200_prof.so")
from pydegensac_b200.scenes import batch_F
P = int(sys.argv[2]) if len(sys.argv) > 2 else 296
plane = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
b1, b2 = batch_F(P, plane_frac=plane)
seeds = np.arange(P, dtype=np.uint64)
L = _cabi.lib()
_cabi.fundamental_batch(b1[:8], b2[:8], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:8])
buf = (ctypes.c_ulonglong * 64)()
L.dgb200_prof_read(buf, 1)
F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
ms = _cabi.last_kernel_ms()
L.dgb200_prof_read(buf, 1)
names = {0: "wave stage A (7pt solve)", 1: "wave stage B (scoring)", 2: "replay total", 3: "  LO total", 4: "    hash",
         5: "  checksample", 6: "  degenerate branch", 7: "    fit_F len<=8", 8: "    fit_F eig"}
print("pairs", P, "kernel ms %.1f" % ms, "-> %.0f pairs/s" % (P / ms * 1e3))
tot = sum(buf[i] for i in (0, 1, 2))
for i in range(9):
    print("%-28s %10.3f Mcycles/pair  %5.1f%%" % (names[i], buf[i] / P / 1e6, 100.0 * buf[i] / max(tot, 1)))
print("waves/pair %.1f  iterations waved/pair %.0f  candidates/pair %.0f  A1 %.3f Mcycles/pair" % (buf[9] / P, buf[12] / P, buf[13] / P, buf[10] / P / 1e6))
print("replayed iterations/pair %.1f  replayed models/pair %.1f  best-sample updates/pair %.2f  LO runs/pair %.2f  re-waves/pair %.2f" % tuple(buf[i] / P for i in (14, 15, 16, 17, 18)))
def per(i, n): return buf[i] / max(buf[n], 1)
print("resid passes/pair %.0f (%.1f kcyc each, %.2f Mcyc/pair)" % (buf[20] / P, per(19, 20) / 1e3, buf[19] / P / 1e6))
print("inlidxs/pair %.0f (%.1f kcyc each, %.2f Mcyc/pair)" % (buf[22] / P, per(21, 22) / 1e3, buf[21] / P / 1e6))
print("randsubset %.2f Mcyc/pair" % (buf[23] / P / 1e6))
print("fits<=8/pair %.0f (%.1f kcyc each)   eig fits/pair %.0f (%.1f kcyc each)" % (buf[27] / P, per(7, 27) / 1e3, buf[28] / P, per(8, 28) / 1e3))
print("rank2 %.2f Mcyc/pair (%.1f kcyc each)   eig solver %.2f Mcyc/pair (%.1f kcyc each)" % (buf[24] / P / 1e6, buf[24] / max(buf[27] + buf[28], 1) / 1e3, buf[29] / P / 1e6, per(29, 28) / 1e3))
print("inner_H/pair %.2f (%.2f Mcyc each, %.2f Mcyc/pair)   rFtH/pair %.2f (%.2f Mcyc each, %.2f Mcyc/pair)" % (buf[30] / P, per(25, 30) / 1e6, buf[25] / P / 1e6, buf[31] / P, per(26, 31) / 1e6, buf[26] / P / 1e6))
print("rFtH: 2-pt hypotheses/pair %.0f  wave time %.2f Mcyc/pair   events/pair %.2f  inner_FH %.2f Mcyc each (%.2f Mcyc/pair)" % (buf[35] / P, buf[32] / P / 1e6, buf[33] / P, per(34, 33) / 1e6, buf[34] / P / 1e6))
print("inner_FH: dual_sample %.2f Mcyc/pair   u2Fit/pair %.2f (%.0f kcyc each, %.2f Mcyc/pair)" % (buf[36] / P / 1e6, buf[38] / P, per(37, 38) / 1e3, buf[37] / P / 1e6))
print("big fits/pair %.1f (%.0f kcyc each, %.2f Mcyc/pair)" % (buf[39] / P, per(40, 39) / 1e3, buf[40] / P / 1e6))
print("fused 8-pt fit: sample+replay %.1f k  gather+weights %.1f k  GJ %.1f k  rank2+store %.1f k  closing barrier %.1f k  (per call)" % tuple(buf[i] / max(buf[27], 1) / 1e3 for i in (41, 42, 43, 44, 45)))
print("rFtH wave parts: swap replay (thread 0) %.2f Mcyc/pair   hypothesis loop %.2f Mcyc/pair" % (buf[47] / P / 1e6, buf[48] / P / 1e6))
