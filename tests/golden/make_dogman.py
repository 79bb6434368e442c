"""Freezes the point set of BASELINE.json configs[0] (examples/simple-example.py:39-53 of the reference: AKAZE
descriptor_type=3, threshold=1e-5 -> BFMatcher knn k=2 -> SNN 0.9 on HPatches v_dogman 1<->6) together with
the reference's answers on it.  Needs cv2 and /root/reference (build container only)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

IMG = "/root/reference/examples/img/v_dogman"
img1 = cv2.cvtColor(cv2.imread(os.path.join(IMG, "1.ppm")), cv2.COLOR_BGR2RGB)
img2 = cv2.cvtColor(cv2.imread(os.path.join(IMG, "6.ppm")), cv2.COLOR_BGR2RGB)
det = cv2.AKAZE_create(descriptor_type=3, threshold=0.00001)
kps1, descs1 = det.detectAndCompute(img1, None)
kps2, descs2 = det.detectAndCompute(img2, None)
bf = cv2.BFMatcher()
matches = bf.knnMatch(descs1, descs2, k=2)
tentatives = [m[0] for m in matches if m[0].distance < 0.9 * m[1].distance]
src = np.float32([kps1[m.queryIdx].pt for m in tentatives]).reshape(-1, 2).astype(np.float64)
dst = np.float32([kps2[m.trainIdx].pt for m in tentatives]).reshape(-1, 2).astype(np.float64)
H_gt = np.loadtxt(os.path.join(IMG, "H_1_6"))
out = dict(src=src, dst=dst, H_gt=H_gt)
for seed in (0, 1):
    Hraw, mask, stats = ref.find_homography_raw(src, dst, 4.0, 0.99, 2000, seed=seed)
    out["H_raw_%d" % seed] = Hraw; out["H_mask_%d" % seed] = mask; out["H_stats_%d" % seed] = stats
    F, fmask, fstats = ref.find_fundamental(src, dst, 0.5, 0.999, 50000, seed=seed)
    out["F_%d" % seed] = F; out["F_mask_%d" % seed] = fmask; out["F_stats_%d" % seed] = fstats
    print(seed, len(src), stats, int(mask.sum()), fstats, int(fmask.sum()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dogman_v1.npz"), **out)
