"""numpy restatements of the steps either side of the RANSAC path -- TEST INFRASTRUCTURE ONLY.

* match_descriptors: what the reference's example pipeline does with cv2.BFMatcher().knnMatch(k=2) + the SNN ratio
  test (examples/simple-example.py:46-53), as a brute-force float32 distance matrix.
* pose_from_fundamental: E = K2^T F K1 -> SVD -> four (R, t) candidates -> cheirality vote (Hartley & Zisserman
  section 9.6; the reference itself stops at F).  tests/ pin it to cv2.recoverPose on scenes with known pose."""
import numpy as np


def match_descriptors(d1, d2, ratio=0.9, mutual=False):
    d1 = np.asarray(d1, np.float32); d2 = np.asarray(d2, np.float32)
    dist = np.zeros((len(d1), len(d2)), np.float32)
    for i in range(len(d1)):            # accumulate in descriptor order, float32, like the kernel
        diff = d2 - d1[i]
        dist[i] = np.einsum("jk,jk->j", diff, diff, dtype=np.float32)
    order = np.argsort(dist, 1, kind="stable")[:, :2]
    nn, b1, b2 = order[:, 0], dist[np.arange(len(d1)), order[:, 0]], dist[np.arange(len(d1)), order[:, 1]]
    ok = b1 < np.float32(ratio * ratio) * b2
    if mutual:
        back = np.argmin(dist, 0)
        ok &= back[nn] == np.arange(len(d1))
    q = np.nonzero(ok)[0]
    return q, nn[q], dist


def pose_from_fundamental(F, K1, K2, p1, p2, mask=None):
    E = K2.T @ F @ K1
    U, S, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    cands = [(U @ W @ Vt, U[:, 2]), (U @ W @ Vt, -U[:, 2]), (U @ W.T @ Vt, U[:, 2]), (U @ W.T @ Vt, -U[:, 2])]
    x1 = np.linalg.solve(K1, np.c_[p1, np.ones(len(p1))].T).T
    x2 = np.linalg.solve(K2, np.c_[p2, np.ones(len(p2))].T).T
    if mask is not None:
        x1, x2 = x1[mask], x2[mask]
    votes = []
    for R, t in cands:
        a1 = x1 @ R.T
        aa = (a1 * a1).sum(1); bb = (x2 * x2).sum(1); ab = (a1 * x2).sum(1); at = a1 @ t; bt = x2 @ t
        det = aa * bb - ab * ab
        l1 = (-at * bb + ab * bt) / det
        l2 = (aa * bt - ab * at) / det
        votes.append(int(((l1 > 0) & (l2 > 0)).sum()))
    best = int(np.argmax(votes))
    return cands[best][0], cands[best][1], votes[best]
