import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import batch_F
P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
b1, b2 = batch_F(P)
F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, np.arange(P, dtype=np.uint64))
print("done", m.sum(1).mean())
