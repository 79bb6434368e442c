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
buf = (ctypes.c_ulonglong * 32)()
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
