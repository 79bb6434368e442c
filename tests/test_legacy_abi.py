"""The reference's own C entry points served by the engine (include/degensac_legacy.h, SURVEY.md section 8(b) proposal 4)."""
import ctypes
import os

import numpy as np
import pytest

from pydegensac_b200.scenes import scene_F, scene_H, scene_F_laf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pydegensac_b200", "libdegensac_b200_legacy.so")
DECLARED = ["exp_ransacFcustomLAF", "exp_ransacHcustomLAF", "FDs", "exFDs", "FDsidx", "FDsSym", "exFDsSym", "FDsSymidx",
            "HDs", "HDsi", "HDsidx", "HDsSymMaxSq", "HDsiSymMaxSq", "HDsSymMaxSqidx", "HDsSymMax", "HDsiSymMax",
            "HDsSymMaxidx", "HDsSymSumSq", "HDsiSymSumSq", "HDsSymSumSqidx", "HDsSymSum", "HDsiSymSum", "HDsSymSumidx"]


class Score(ctypes.Structure):
    _fields_ = [("I", ctypes.c_uint), ("J", ctypes.c_double), ("Is", ctypes.c_uint), ("Ilafs", ctypes.c_uint)]


def test_legacy_library_exports_the_reference_symbols():
    L = ctypes.CDLL(LIB)
    hdr = open(os.path.join(ROOT, "include", "degensac_legacy.h")).read()
    for name in DECLARED:
        assert name + "(" in hdr.replace(" (", "("), name
        getattr(L, name)


def _u6(p1, p2):
    n = len(p1)
    u = np.ones((n, 6))
    u[:, 0:2] = p1[:, :2]; u[:, 3:5] = p2[:, :2]
    return np.ascontiguousarray(u)


def _helpers(p1, p2):
    """u_1 / u_2 of the binding layer (bindings.cpp:355-385): p1 = x + (a12, a22), p2 = x + (a11, a21)."""
    n = len(p1)
    u1 = np.ones((n, 6)); u2 = np.ones((n, 6))
    for img, P in ((0, p1), (3, p2)):
        u1[:, img] = P[:, 0] + P[:, 3]; u1[:, img + 1] = P[:, 1] + P[:, 5]
        u2[:, img] = P[:, 0] + P[:, 2]; u2[:, img + 1] = P[:, 1] + P[:, 4]
    return np.ascontiguousarray(u1), np.ascontiguousarray(u2)


@pytest.mark.gpu
def test_legacy_entry_points_equal_the_abi_and_the_reference(monkeypatch, ref_oracle):
    from pydegensac_b200 import _cabi
    L = ctypes.CDLL(LIB)
    dp = ctypes.POINTER(ctypes.c_double)
    vp = ctypes.c_void_p
    fn = lambda name: ctypes.cast(getattr(L, name), vp)
    L.exp_ransacFcustomLAF.restype = ctypes.c_int
    L.exp_ransacFcustomLAF.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                       dp, ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                       ctypes.c_uint, ctypes.POINTER(dp), dp, ctypes.POINTER(ctypes.c_int), vp, vp, vp,
                                       ctypes.c_double, ctypes.c_int]
    L.exp_ransacHcustomLAF.restype = Score
    L.exp_ransacHcustomLAF.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                       dp, ctypes.POINTER(ctypes.c_ubyte), ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                       ctypes.c_int, ctypes.c_uint, ctypes.POINTER(dp), vp, vp, vp, ctypes.c_double]
    monkeypatch.setenv("DGB200_LEGACY_SEED", "77")
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [vp]
    # ---- F, both metrics, with and without the LAF helper arrays, exactly as bindings.cpp:420-435 calls the core
    for metric, names in ((0, ("exFDs", "FDs", "FDsidx")), (1, ("exFDsSym", "FDsSym", "FDsSymidx"))):
        for laf in (False, True):
            if laf:
                p1, p2, _ = scene_F_laf(600, 0.5, 4)
            else:
                p1, p2, _ = scene_F(900, 0.4, 3, 0.5)
            n = len(p1)
            u = _u6(p1, p2)
            u1, u2 = _helpers(p1, p2) if laf else (u, u)
            px = 1.0
            F = np.zeros(9); inl = np.zeros(n, np.uint8); data_out = np.zeros(18 * n, np.int32)
            resids = dp(); Ih = ctypes.c_int(0); Hbest = np.zeros(9)
            I = L.exp_ransacFcustomLAF(u.ctypes.data_as(dp), u1.ctypes.data_as(dp), u2.ctypes.data_as(dp), n, px * px,
                                       3.0 if laf else -1.0, 0.999, 3000, F.ctypes.data_as(dp),
                                       inl.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                       data_out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 1, 0, ctypes.byref(resids),
                                       Hbest.ctypes.data_as(dp), ctypes.byref(Ih), fn(names[0]), fn(names[1]), fn(names[2]),
                                       3.0 * px * px, 1)
            libc.free(ctypes.cast(resids, vp))
            Fa, ma, sa = _cabi.fundamental_batch(p1, p2, px, 0.999, 3000, metric, True, 3.0 if laf else 0.0, True, [77])
            assert I == sa[0][3] and data_out[0] == sa[0][0] and data_out[1] == sa[0][1]
            assert np.array_equal(inl.astype(bool), ma[0])
            assert np.allclose(F.reshape(3, 3), Fa[0], rtol=1e-9, atol=1e-12)   # (helper points are re-derived: 1e-16 level)
            r = ref_oracle.find_fundamental(p1, p2, px, 0.999, 3000, error_type=metric, laf_coef=3.0 if laf else 0.0, seed=77)
            assert np.array_equal(r[1], inl.astype(bool))
    # ---- H, the five metrics with the binding's threshold table (bindings.cpp:64-107)
    p1, p2, _ = scene_H(1500, 500, 6)
    n = len(p1); u = _u6(p1, p2); px = 3.0
    table = {0: ("HDs", "HDsi", "HDsidx", px * px, 3 * px), 1: ("HDsSymMaxSq", "HDsiSymMaxSq", "HDsSymMaxSqidx", px * px, 0.0),
             2: ("HDsSymMax", "HDsiSymMax", "HDsSymMaxidx", px, 0.0), 3: ("HDsSymSumSq", "HDsiSymSumSq", "HDsSymSumSqidx", px * px, 3 * px),
             4: ("HDsSymSum", "HDsiSymSum", "HDsSymSumidx", px, 3 * px)}
    for metric, (a, b, c, th, symth) in table.items():
        H = np.zeros(9); inl = np.zeros(n, np.uint8); data_out = np.zeros(18 * n, np.int32); resids = dp()
        S = L.exp_ransacHcustomLAF(u.ctypes.data_as(dp), u.ctypes.data_as(dp), u.ctypes.data_as(dp), n, th, -1.0, 0.999, 3000,
                                   H.ctypes.data_as(dp), inl.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), 4,
                                   data_out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 1, 0, ctypes.byref(resids), fn(a), fn(b),
                                   fn(c), symth)
        libc.free(ctypes.cast(resids, vp))
        Ha, ma, sa = _cabi.homography_batch(p1, p2, px, 0.999, 3000, metric, True, 0.0, [77])
        assert S.I == sa[0][3] and np.array_equal(inl.astype(bool), ma[0]) and np.array_equal(H.reshape(3, 3), Ha[0])
        r = ref_oracle.find_homography_raw(p1, p2, px, 0.999, 3000, error_type=metric, seed=77)
        assert np.array_equal(r[1], inl.astype(bool))
    # an argument combination the binding never produces is refused, the model stays zero
    F = np.ones(9); inl = np.zeros(n, np.uint8); resids = dp(); Ih = ctypes.c_int(0)
    I = L.exp_ransacFcustomLAF(u.ctypes.data_as(dp), u.ctypes.data_as(dp), u.ctypes.data_as(dp), n, 1.0, -1.0, 0.99, 100,
                               F.ctypes.data_as(dp), inl.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), None, 1, 0,
                               ctypes.byref(resids), None, ctypes.byref(Ih), fn("exFDs"), fn("FDsSym"), fn("FDsidx"), 3.0, 1)
    libc.free(ctypes.cast(resids, vp))
    assert I == 0 and np.abs(F).sum() == 0
