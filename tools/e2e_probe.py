"""Where the host-buffer API spends its time: wall clock vs kernel time of dgb200_find_fundamental_batch at P pairs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pydegensac_b200 import _cabi
from pydegensac_b200.scenes import batch_F
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b1, b2 = batch_F(P)
h1 = torch.from_numpy(b1).pin_memory().numpy(); h2 = torch.from_numpy(b2).pin_memory().numpy()
seeds = np.arange(P, dtype=np.uint64)
for name, (a1, a2) in {"pinned": (h1, h2), "pageable": (b1, b2)}.items():
    _cabi.fundamental_batch(a1[:64], a2[:64], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[:64])
    for rep in range(3):
        t = time.perf_counter()
        _cabi.fundamental_batch(a1, a2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
        w = time.perf_counter() - t
        print(name, "wall %.1f ms, kernel %.1f ms, launches %d" % (w * 1e3, _cabi.last_kernel_ms(), _cabi.kernel_launches()), flush=True)
import pydegensac_b200 as pdg
for rep in range(3):
    t = time.perf_counter()
    F, m = pdg.findFundamentalMatrixBatch(h1, h2, 1.0, 0.9999, 10000, seeds=seeds)
    w = time.perf_counter() - t
    print("public API (pinned) wall %.1f ms, kernel %.1f ms" % (w * 1e3, _cabi.last_kernel_ms()), flush=True)
