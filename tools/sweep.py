import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
if os.environ.get('DGLIB'): _cabi._LIBPATH = os.path.abspath(os.environ['DGLIB'])
from pydegensac_b200.scenes import batch_F
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
b1, b2 = batch_F(P)
seeds = np.arange(P, dtype=np.uint64)
_cabi.fundamental_batch(b1[:64], b2[:64], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:64])
for n in [int(x) for x in sys.argv[2:]] or [P]:
    F, m, s = _cabi.fundamental_batch(b1[:n], b2[:n], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:n])
    ms = _cabi.last_kernel_ms()
    print(os.environ.get("DGLIB"), "threads", os.environ.get("DGB200_THREADS"), "tile", os.environ.get("DGB200_SMEM_TILE"), "pairs", n, "kernel %.1f ms -> %.0f pairs/s" % (ms, n / ms * 1e3), "inl %.1f" % m.sum(1).mean(), flush=True)
