"""The steps either side of the RANSAC path: GPU descriptor matcher, device-tensor
entry points, pose from F, the QR null-space solver."""
import numpy as np
import pytest

from pydegensac_b200.scenes import scene_F, scene_H, _K


def _scene_pose(n, seed):
    """Two views of random 3-D points with a KNOWN relative pose."""
    from pydegensac_b200.scenes import _rot
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(4, 10, n)]
    R = _rot("z", 0.03) @ _rot("y", -0.2) @ _rot("x", 0.05)
    t = np.array([1.0, 0.1, 0.2])
    x1 = (_K @ X.T).T
    x2 = (_K @ (R @ X.T + t[:, None])).T
    return x1[:, :2] / x1[:, 2:3], x2[:, :2] / x2[:, 2:3], R, t / np.linalg.norm(t)


def test_pose_oracle_matches_cv2_and_ground_truth():
    """Pins oracle/frontend_np.pose_from_fundamental (the GPU kernel's checker) to OpenCV and to the planted pose."""
    cv2 = pytest.importorskip("cv2")
    from oracle.frontend_np import pose_from_fundamental
    for seed in range(4):
        p1, p2, R, t = _scene_pose(300, seed)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        F = np.linalg.inv(_K).T @ (tx @ R) @ np.linalg.inv(_K)
        Ro, to, good = pose_from_fundamental(F, _K, _K, p1, p2)
        assert good == 300
        assert np.allclose(Ro, R, atol=1e-8) and np.allclose(to, t, atol=1e-8)
        E = _K.T @ F @ _K
        _, Rc, tc, _ = cv2.recoverPose(E, p1, p2, _K)
        assert np.allclose(Rc, Ro, atol=1e-6) and np.allclose(tc.ravel(), to, atol=1e-6)


def test_qr_nullspace_matches_lapack_dgeqp3():
    """la.h nullspace_qr7x9 (restated dgeqp3 kernel + the reference's back substitution, Ftools.c:594-668) against the
    same LAPACK routine reached through scipy (oracle/qr7x9_np.py)."""
    import ctypes
    from tests.hostemu import emu
    from oracle.qr7x9_np import nullspace_qr7x9
    E = emu.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    E.emu_nullspace_qr7x9.argtypes = [dp, dp]
    rng = np.random.default_rng(0)
    for t in range(200):
        if t < 100:
            A = rng.normal(size=(7, 9))
        else:
            p1, p2, _ = scene_F(50, 1.0, t)
            idx = rng.choice(50, 7, replace=False)
            A = np.stack([[p2[i, 0] * p1[i, 0], p2[i, 0] * p1[i, 1], p2[i, 0], p2[i, 1] * p1[i, 0], p2[i, 1] * p1[i, 1],
                           p2[i, 1], p1[i, 0], p1[i, 1], 1.0] for i in idx])
        A = np.ascontiguousarray(A)
        Nb = np.zeros(18)
        ra, Na = nullspace_qr7x9(A)
        rb = E.emu_nullspace_qr7x9(A.ctypes.data_as(dp), Nb.ctypes.data_as(dp))
        assert ra == rb
        assert np.abs(Na.ravel() - Nb).max() <= 1e-9 * max(1.0, np.abs(Na).max())


@pytest.mark.gpu
def test_gpu_matcher_vs_bruteforce_and_pipeline():
    import torch
    import pydegensac_b200 as pdg
    from pydegensac_b200.matching import match_descriptors
    from oracle.frontend_np import match_descriptors as match_np
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    for (n1, n2, D, mutual) in [(700, 900, 64, False), (1500, 1200, 128, True), (33, 65, 8, True), (2000, 2000, 64, True)]:
        # scene: a homography-related keypoint set; descriptors = shared random codes + noise for true matches
        n_true = min(n1, n2) // 2
        p1, p2, _ = scene_H(n_true, n_true, n1 + D)
        kp1 = np.r_[p1, rng.uniform(0, 640, (n1 - n_true, 2))]
        kp2 = np.r_[p2, rng.uniform(0, 640, (n2 - n_true, 2))]
        base = rng.normal(size=(n_true, D)).astype(np.float32)
        d1 = np.r_[base + 0.15 * rng.normal(size=(n_true, D)).astype(np.float32), rng.normal(size=(n1 - n_true, D)).astype(np.float32)]
        d2 = np.r_[base + 0.15 * rng.normal(size=(n_true, D)).astype(np.float32), rng.normal(size=(n2 - n_true, D)).astype(np.float32)]
        perm = rng.permutation(n2)
        d2, kp2 = d2[perm], kp2[perm]
        q_ref, t_ref, dist = match_np(d1, d2, 0.9, mutual)
        i1, i2, x1, x2 = match_descriptors(torch.from_numpy(d1).to(dev), torch.from_numpy(d2).to(dev),
                                           torch.from_numpy(kp1).to(dev), torch.from_numpy(kp2).to(dev), 0.9, mutual)
        i1, i2 = i1.cpu().numpy(), i2.cpu().numpy()
        # identical match lists, except ratio decisions within float32 rounding of the threshold
        s = np.sort(dist, 1)
        margin = np.abs(s[:, 0] - np.float32(0.81) * s[:, 1]) < 1e-4 * s[:, 1]
        a = set(zip(q_ref.tolist(), t_ref.tolist())); b = set(zip(i1.tolist(), i2.tolist()))
        for (q, t) in a ^ b:
            assert margin[q], "match (%d,%d) differs away from the ratio threshold" % (q, t)
        assert np.all(np.diff(i1) > 0)
        assert np.array_equal(x1.cpu().numpy(), kp1[i1]) and np.array_equal(x2.cpu().numpy(), kp2[i2])
        assert len(b) >= 0.8 * n_true
    # pipeline: matcher output -> device-resident RANSAC, nothing on the host in between
    H, mask = pdg.findHomographyBatch(x1[None], x2[None], 3.0, 0.999, 2000, seeds=[1])
    assert isinstance(H, torch.Tensor) and H.is_cuda and mask.dtype == torch.bool
    Hn, mn = pdg.findHomographyBatch(x1.cpu().numpy()[None], x2.cpu().numpy()[None], 3.0, 0.999, 2000, seeds=[1])
    assert np.array_equal(mask.cpu().numpy(), mn) and np.allclose(H.cpu().numpy(), Hn, rtol=1e-12, atol=1e-12)
    assert mn.sum() >= 0.7 * len(i1)


@pytest.mark.gpu
def test_gpu_tensor_entry_points_equal_host_entry_points():
    import torch
    import pydegensac_b200 as pdg
    from pydegensac_b200.scenes import batch_F
    dev = torch.device("cuda:0")
    b1, b2 = batch_F(24, 1500, 0.4, seed0=11)
    seeds = np.arange(24, dtype=np.uint64) + 5
    Fh, mh, sh = pdg.findFundamentalMatrixBatch(b1, b2, 1.0, 0.999, 3000, seeds=seeds, return_stats=True)
    Ft, mt, st = pdg.findFundamentalMatrixBatch(torch.from_numpy(b1).to(dev), torch.from_numpy(b2).to(dev), 1.0, 0.999,
                                                3000, seeds=seeds, return_stats=True)
    assert Ft.is_cuda and np.array_equal(Ft.cpu().numpy(), Fh) and np.array_equal(mt.cpu().numpy(), mh)
    assert np.array_equal(st.cpu().numpy(), sh)
    # float32 tensors and DLPack exporters are accepted too
    class Exporter:
        def __init__(self, t): self.t = t
        def __dlpack__(self, stream=None): return self.t.__dlpack__()
        def __dlpack_device__(self): return self.t.__dlpack_device__()
    Fd, md = pdg.findFundamentalMatrixBatch(Exporter(torch.from_numpy(b1).to(dev)), Exporter(torch.from_numpy(b2).to(dev)),
                                            1.0, 0.999, 3000, seeds=seeds)
    assert np.array_equal(Fd.cpu().numpy(), Fh)


@pytest.mark.gpu
def test_gpu_pose_and_qr_kernels():
    import ctypes
    import torch
    from pydegensac_b200 import _cabi
    from pydegensac_b200.matching import pose_from_fundamental
    from oracle.frontend_np import pose_from_fundamental as pose_np
    from oracle.qr7x9_np import nullspace_qr7x9
    dev = torch.device("cuda:0")
    Fs, P1, P2, Rs, ts = [], [], [], [], []
    for seed in range(6):
        p1, p2, R, t = _scene_pose(400, seed)
        p1 = p1 + np.random.default_rng(seed).normal(0, 0.2, p1.shape)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Fs.append(np.linalg.inv(_K).T @ (tx @ R) @ np.linalg.inv(_K)); P1.append(p1); P2.append(p2); Rs.append(R); ts.append(t)
    mask = np.ones((6, 400), bool); mask[:, ::7] = False
    K = torch.from_numpy(_K).to(dev)
    R, t, good = pose_from_fundamental(torch.from_numpy(np.stack(Fs)).to(dev), K, K, torch.from_numpy(np.stack(P1)).to(dev),
                                       torch.from_numpy(np.stack(P2)).to(dev), torch.from_numpy(mask).to(dev))
    R, t, good = R.cpu().numpy(), t.cpu().numpy(), good.cpu().numpy()
    for i in range(6):
        Ro, to, go = pose_np(Fs[i], _K, _K, P1[i], P2[i], mask[i])
        assert np.allclose(R[i], Ro, atol=1e-9) and np.allclose(t[i], to, atol=1e-9) and good[i] == go
        assert np.allclose(R[i], Rs[i], atol=1e-6) and np.allclose(t[i], ts[i], atol=1e-6)
    rng = np.random.default_rng(1)
    A = rng.normal(size=(500, 7, 9))
    dA = torch.from_numpy(A).to(dev); dN = torch.zeros((500, 18), dtype=torch.float64, device=dev)
    rc = torch.zeros(500, dtype=torch.int32, device=dev)
    vp = ctypes.c_void_p
    assert _cabi.lib().dgb200_nullspace_qr7x9_batch_dev(vp(dA.data_ptr()), vp(dN.data_ptr()), vp(rc.data_ptr()), 500, vp(0)) == 0
    torch.cuda.synchronize()
    N = dN.cpu().numpy()
    for i in range(0, 500, 7):
        r, Nr = nullspace_qr7x9(A[i])
        assert r == 0 and np.abs(Nr.ravel() - N[i]).max() <= 1e-9 * max(1.0, np.abs(Nr).max())
