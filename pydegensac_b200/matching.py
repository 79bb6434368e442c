"""GPU front-end of the RANSAC path: descriptor matching on the device.

Replaces the host-side matching of the reference's pipeline -- `cv2.BFMatcher().knnMatch(descs1, descs2, k=2)` followed
by the SNN ratio test `m.distance < 0.9 * n.distance` (examples/simple-example.py:46-53) -- with a CUDA kernel whose
output, two [N,2] float64 CUDA tensors, is what `findFundamentalMatrixBatch / findHomographyBatch` take: keypoints and
descriptors go in, models and masks come out, nothing visits the host in between.

    idx1, idx2, pts1, pts2 = match_descriptors(desc1, desc2, kp1, kp2, ratio=0.9)
    F, mask = pydegensac_b200.findFundamentalMatrixBatch(pts1[None], pts2[None], 0.5, 0.999, 50000)
"""
import ctypes

from . import _cabi


def match_descriptors(desc1, desc2, kp1=None, kp2=None, ratio=0.9, mutual=False):
    """Two-nearest-neighbour L2 matching + SNN ratio test (+ optional mutual nearest-neighbour check).

    desc1 [n1,D], desc2 [n2,D]: CUDA tensors (any float dtype; computed in float32), D a multiple of 4, <= 256.
    kp1 [n1,>=2], kp2 [n2,>=2]: optional keypoint coordinate tensors (first two columns x, y; six columns carry the
    local affine shape through for the LAF check).
    Returns (idx1, idx2) int64 CUDA tensors of the accepted matches in ascending query order, plus (pts1, pts2) float64
    [N,2] (or [N,6]) CUDA tensors when keypoints were given."""
    import torch
    d1 = desc1.to(torch.float32).contiguous()
    d2 = desc2.to(torch.float32).contiguous()
    if d1.dim() != 2 or d2.dim() != 2 or d1.shape[1] != d2.shape[1] or not d1.is_cuda or d1.device != d2.device:
        raise ValueError("expected two CUDA descriptor tensors [n1,D], [n2,D] on the same device")
    n1, D = d1.shape
    n2 = d2.shape[0]
    dev = d1.device
    L = _cabi.lib()
    with torch.cuda.device(dev):
        mq = torch.empty(n1, dtype=torch.int32, device=dev)
        mt = torch.empty(n1, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(int(L.dgb200_match_workspace_bytes(n1, n2)), dtype=torch.uint8, device=dev)
        k1 = k2 = x1 = x2 = None
        kp_dim = out_dim = 0
        if kp1 is not None:
            k1 = kp1.to(device=dev, dtype=torch.float64).contiguous()
            k2 = kp2.to(device=dev, dtype=torch.float64).contiguous()
            if k1.shape[0] != n1 or k2.shape[0] != n2 or k1.shape[1] != k2.shape[1] or k1.shape[1] < 2:
                raise ValueError("keypoint tensors must be [n1,k], [n2,k] with k >= 2")
            kp_dim = int(k1.shape[1])
            out_dim = 6 if kp_dim >= 6 else 2
            x1 = torch.empty((n1, out_dim), dtype=torch.float64, device=dev)
            x2 = torch.empty((n1, out_dim), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        vp = ctypes.c_void_p
        rc = L.dgb200_match_descriptors_dev(vp(d1.data_ptr()), n1, vp(d2.data_ptr()), n2, D, ctypes.c_float(ratio),
                                            int(bool(mutual)), vp(k1.data_ptr()) if k1 is not None else None,
                                            vp(k2.data_ptr()) if k2 is not None else None, kp_dim, vp(mq.data_ptr()),
                                            vp(mt.data_ptr()), vp(x1.data_ptr()) if x1 is not None else None,
                                            vp(x2.data_ptr()) if x2 is not None else None, out_dim, n1,
                                            vp(cnt.data_ptr()), vp(ws.data_ptr()), vp(stream))
        if rc != 0:
            raise ValueError(L.dgb200_frontend_last_error().decode())
        n = int(cnt.item())          # the one host read of the front-end: how many tentatives there are
    out = (mq[:n].long(), mt[:n].long())
    if x1 is not None:
        out = out + (x1[:n], x2[:n])
    return out


def pose_from_fundamental(F, K1, K2, pts1, pts2, mask=None):
    """Relative pose from fundamental matrices: E = K2^T F K1, SVD, cheirality vote over the inliers.

    F [P,3,3], pts [P,N,2] (float64 CUDA tensors), K1/K2 [3,3] or [P,3,3], mask [P,N] bool/uint8 or None.
    Returns R [P,3,3], t [P,3] (unit length; x2 ~ R x1 + t), good [P] int32 (correspondences in front of both cameras)."""
    import torch
    dev = F.device
    Fc = F.to(torch.float64).contiguous().view(-1, 9)
    P = Fc.shape[0]
    p1 = pts1.to(torch.float64).contiguous()
    p2 = pts2.to(torch.float64).contiguous()
    if p1.dim() == 2:
        p1, p2 = p1[None], p2[None]
    N, dim = int(p1.shape[1]), int(p1.shape[2])
    K1c = K1.to(device=dev, dtype=torch.float64).contiguous()
    K2c = K2.to(device=dev, dtype=torch.float64).contiguous()
    per_pair = 1 if K1c.dim() == 3 else 0
    m = None
    if mask is not None:
        m = mask.to(device=dev, dtype=torch.uint8).contiguous()
    L = _cabi.lib()
    with torch.cuda.device(dev):
        R = torch.empty((P, 3, 3), dtype=torch.float64, device=dev)
        t = torch.empty((P, 3), dtype=torch.float64, device=dev)
        good = torch.empty(P, dtype=torch.int32, device=dev)
        vp = ctypes.c_void_p
        rc = L.dgb200_pose_from_fundamental_batch_dev(vp(Fc.data_ptr()), vp(K1c.data_ptr()), vp(K2c.data_ptr()), per_pair,
                                                      vp(p1.data_ptr()), vp(p2.data_ptr()),
                                                      vp(m.data_ptr()) if m is not None else None, P, N, dim,
                                                      vp(R.data_ptr()), vp(t.data_ptr()), vp(good.data_ptr()),
                                                      vp(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise ValueError(L.dgb200_frontend_last_error().decode())
    return R, t, good
