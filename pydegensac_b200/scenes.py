"""Synthetic two-view scenes used by the benchmark and the parity tests.

These reproduce the configurations BASELINE.json names; they are data generators only.
"""
import numpy as np

_K = np.array([[800.0, 0.0, 320.0], [0.0, 800.0, 240.0], [0.0, 0.0, 1.0]])


def _rot(ax, a):
    c, s = np.cos(a), np.sin(a)
    if ax == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)
    if ax == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def scene_F(n=2000, inlier_ratio=0.30, seed=0, plane_frac=0.0, noise=0.3):
    """Scene F(N, rho, seed, pi): two pinhole views of random 3-D points + uniform outliers.

    Returns (pts1 [N,2], pts2 [N,2], is_inlier [N]) float64; inliers first, then outliers.
    """
    rng = np.random.default_rng(seed)
    n_in = int(n * inlier_ratio)
    X = np.empty((n_in, 3))
    X[:, 0] = rng.uniform(-2, 2, n_in)
    X[:, 1] = rng.uniform(-2, 2, n_in)
    X[:, 2] = rng.uniform(4, 10, n_in)
    n_pl = int(n_in * plane_frac)
    if n_pl > 0:
        X[:n_pl, 2] = 6.0 + 0.1 * X[:n_pl, 0]
    R = _rot("z", 0.03) @ _rot("y", -0.2) @ _rot("x", 0.05)
    t = np.array([1.0, 0.1, 0.2])
    x1 = (_K @ X.T).T
    x2 = (_K @ (R @ X.T + t[:, None])).T
    p1 = x1[:, :2] / x1[:, 2:3] + rng.normal(0, noise, (n_in, 2))
    p2 = x2[:, :2] / x2[:, 2:3] + rng.normal(0, noise, (n_in, 2))
    n_out = n - n_in
    o1 = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 480, n_out)], 1)
    o2 = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 480, n_out)], 1)
    pts1 = np.concatenate([p1, o1], 0)
    pts2 = np.concatenate([p2, o2], 0)
    gt = np.zeros(n, dtype=bool)
    gt[:n_in] = True
    return np.ascontiguousarray(pts1), np.ascontiguousarray(pts2), gt


H_GT = np.array([[1.1, 0.05, 10.0], [-0.03, 0.95, -5.0], [1e-4, -5e-5, 1.0]])


def scene_H(n=5000, n_in=1500, seed=0, noise=0.5):
    """H scene of BASELINE config 3: b = proj(H_GT a) + N(0, noise) on inliers, uniform outliers."""
    rng = np.random.default_rng(seed)
    a = np.stack([rng.uniform(0, 640, n_in), rng.uniform(0, 640, n_in)], 1)
    ah = np.concatenate([a, np.ones((n_in, 1))], 1) @ H_GT.T
    b = ah[:, :2] / ah[:, 2:3] + rng.normal(0, noise, (n_in, 2))
    n_out = n - n_in
    o1 = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 640, n_out)], 1)
    o2 = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 640, n_out)], 1)
    pts1 = np.concatenate([a, o1], 0)
    pts2 = np.concatenate([b, o2], 0)
    gt = np.zeros(n, dtype=bool)
    gt[:n_in] = True
    return np.ascontiguousarray(pts1), np.ascontiguousarray(pts2), gt


def batch_F(n_pairs, n=2000, inlier_ratio=0.30, seed0=0, plane_frac=0.0):
    """Config 5: n_pairs scenes F(n, rho, seed=s) for s = seed0 .. seed0+n_pairs-1, stacked [P,N,2]."""
    p1 = np.empty((n_pairs, n, 2))
    p2 = np.empty((n_pairs, n, 2))
    for s in range(n_pairs):
        a, b, _ = scene_F(n, inlier_ratio, seed0 + s, plane_frac)
        p1[s], p2[s] = a, b
    return p1, p2


def scene_F_laf(n=600, inlier_ratio=0.5, seed=0, jitter=0.6, plane_frac=0.0):
    """scene_F plus local affine frames: [N,6] rows (x, y, a11, a12, a21, a22) for the LAF-consistency gate.
    Shapes are ~6 px, equal up to `jitter` px between the two images on inliers, unrelated on outliers."""
    p1, p2, gt = scene_F(n, inlier_ratio, seed, plane_frac)
    rng = np.random.default_rng(seed + 1000)
    A1 = np.tile(np.array([6.0, 0.0, 0.0, 6.0]), (n, 1)) + rng.normal(0, 1.0, (n, 4))
    A2 = A1 + rng.normal(0, jitter, (n, 4))
    A2[~gt] = rng.normal(0, 6, ((~gt).sum(), 4))
    return np.hstack([p1, A1]), np.hstack([p2, A2]), gt


def scene_H_laf(n=800, n_in=400, seed=0, jitter=0.6):
    """scene_H plus local affine frames ([N,6] rows), as scene_F_laf."""
    p1, p2, gt = scene_H(n, n_in, seed)
    rng = np.random.default_rng(seed + 1000)
    A1 = np.tile(np.array([6.0, 0.0, 0.0, 6.0]), (n, 1)) + rng.normal(0, 1.0, (n, 4))
    A2 = A1 + rng.normal(0, jitter, (n, 4))
    A2[~gt] = rng.normal(0, 6, ((~gt).sum(), 4))
    return np.hstack([p1, A1]), np.hstack([p2, A2]), gt


# ---- correspondences of local elliptical features under a planted homography (ransacH2el tests).  Rows u10 =
# (x', y', a', b', c', x, y, a, b, c) as ranH2el.h:4 documents: image 1 first; each frame the lower-triangular
# affinity [a 0; b c] at (x, y).


def _lower(A):
    """A R = L with R a rotation and L lower triangular with positive diagonal (same ellipse A * unit circle)."""
    q, r = np.linalg.qr(A.T)       # A.T = q r  ->  A = r.T q.T
    L = r.T
    s = np.sign(np.diag(L)); s[s == 0] = 1
    return L * s[None, :]


def scene_H2el(n, inlier_frac, seed, noise=0.3, size=800.0, frame_noise=0.01):
    rng = np.random.RandomState(seed)
    H = np.eye(3) + rng.uniform(-0.15, 0.15, (3, 3)) * np.array([[1, 1, 100], [1, 1, 100], [2e-4, 2e-4, 0]])
    H[2, 2] = 1.0
    u = np.zeros((n, 10))
    n_in = int(round(n * inlier_frac))
    for i in range(n):
        x, y = rng.uniform(50, size - 50, 2)
        A = _lower(np.array([[rng.uniform(5, 20), 0], [rng.uniform(-5, 5), rng.uniform(5, 20)]]))
        if i < n_in:
            w = H @ np.array([x, y, 1.0])
            xp, yp = w[0] / w[2], w[1] / w[2]
            # Jacobian of the projective map at (x, y)
            J = (H[:2, :2] - np.outer([xp, yp], H[2, :2])) / w[2]
            Ap = _lower(J @ A)
            if noise > 0:
                xp += rng.normal(0, noise); yp += rng.normal(0, noise)
            Ap = _lower(Ap * (1 + rng.normal(0, frame_noise, (2, 2)))) if frame_noise > 0 else Ap
        else:
            xp, yp = rng.uniform(50, size - 50, 2)
            Ap = _lower(np.array([[rng.uniform(5, 20), 0], [rng.uniform(-5, 5), rng.uniform(5, 20)]]))
        u[i] = [xp, yp, Ap[0, 0], Ap[1, 0], Ap[1, 1], x, y, A[0, 0], A[1, 0], A[1, 1]]
    perm = rng.permutation(n)
    truth = np.zeros(n, dtype=bool); truth[:n_in] = True
    return u[perm], truth[perm], H
