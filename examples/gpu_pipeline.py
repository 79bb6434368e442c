#!/usr/bin/env python
"""The reference's examples/simple-example.py pipeline with every step after feature extraction on the GPU:

    descriptors + keypoints (CUDA tensors)  ->  2-NN matching + SNN ratio test   (pydegensac_b200.matching, CUDA)
                                            ->  findHomography / findFundamentalMatrix (device-resident tensors)
                                            ->  relative pose from F                      (pydegensac_b200.matching)

The reference does the matching with cv2.BFMatcher on the host (simple-example.py:46-53) and copies the tentative
correspondences into numpy arrays; here nothing visits the host between the descriptors and the inlier mask.
Feature extraction itself (AKAZE/SIFT) is outside the path: this script synthesises keypoints and descriptors.

    python examples/gpu_pipeline.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydegensac_b200 as pydegensac
from pydegensac_b200.matching import match_descriptors, pose_from_fundamental
from pydegensac_b200.scenes import scene_F, _K


def synth_features(n_true=1200, n_clutter=800, D=64, seed=0):
    """Two views of one scene: keypoints of scene_F + descriptors that agree (up to noise) on true matches."""
    rng = np.random.default_rng(seed)
    p1, p2, _ = scene_F(n_true, 1.0, seed)
    kp1 = np.r_[p1, rng.uniform(0, 640, (n_clutter, 2))]
    kp2 = np.r_[p2, rng.uniform(0, 640, (n_clutter, 2))]
    base = rng.normal(size=(n_true, D)).astype(np.float32)
    d1 = np.r_[base + 0.2 * rng.normal(size=base.shape).astype(np.float32), rng.normal(size=(n_clutter, D)).astype(np.float32)]
    d2 = np.r_[base + 0.2 * rng.normal(size=base.shape).astype(np.float32), rng.normal(size=(n_clutter, D)).astype(np.float32)]
    perm = rng.permutation(len(kp2))
    return kp1, d1, kp2[perm], d2[perm]


def main():
    dev = torch.device("cuda:0")
    kp1, d1, kp2, d2 = (torch.from_numpy(a).to(dev) for a in synth_features())
    # SNN ratio test 0.9 as in the reference's example; tentatives stay on the device
    idx1, idx2, pts1, pts2 = match_descriptors(d1, d2, kp1, kp2, ratio=0.9, mutual=True)
    print("tentative correspondences:", len(idx1))
    F, mask = pydegensac.findFundamentalMatrixBatch(pts1[None], pts2[None], 0.5, 0.999, 50000, seeds=[0])
    print("pydegensac_b200 found %d inliers" % int(mask.sum()))
    K = torch.from_numpy(_K).to(dev)
    R, t, good = pose_from_fundamental(F, K, K, pts1[None], pts2[None], mask)
    print("R =\n", R[0].cpu().numpy(), "\nt =", t[0].cpu().numpy(), " (%d correspondences in front of both cameras)" % int(good[0]))


if __name__ == "__main__":
    main()
