import sys, os
sys.path.insert(0, '.')
import numpy as np
from pydegensac_b200 import _cabi
if os.environ.get('DGLIB'): _cabi._LIBPATH = os.path.abspath(os.environ['DGLIB'])
from pydegensac_b200.scenes import scene_H
G = np.load('tests/golden/golden_v1.npz')
p1, p2, _ = scene_H(5000, 1500, 0)
for i, et, seed in ((12, 0, 0), (13, 1, 1)):
    M, m, s = _cabi.homography_batch(p1, p2, 3.0, 0.999, 10000, et, True, 0.0, [seed])
    print(os.environ.get("DGLIB"), os.environ.get("DGB200_THREADS"), i, "stats", s[0], "gold", G["stats_%d" % i], "maskdiff", (m[0] != G["mask_%d" % i]).sum(), "model", np.abs(M[0]/np.linalg.norm(M[0])).round(4)[0], np.abs(G["model_%d"%i]/np.linalg.norm(G["model_%d"%i])).round(4)[0])
