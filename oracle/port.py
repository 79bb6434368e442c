"""ctypes loader for oracle/port/libdegensac_port.so -- the plain-C CPU restatement of the hot path.
TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "port")
_SO = os.path.join(_DIR, "libdegensac_port.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_DIR, "degensac_port.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def find_fundamental(pts1, pts2, px_th=0.5, conf=0.9999, max_iters=100000, error_type=0, sym_check=True, laf_coef=0.0,
                     degen_check=True, seed=0, final_lsq=False):
    lib().port_set_final_lsq(int(bool(final_lsq)))
    p1 = np.ascontiguousarray(pts1, dtype=np.float64); p2 = np.ascontiguousarray(pts2, dtype=np.float64)
    n, dim = p1.shape
    F = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); stats = np.zeros(4, dtype=np.int32)
    rc = lib().port_find_fundamental(_dp(p1), _dp(p2), n, dim, ctypes.c_double(px_th), ctypes.c_double(conf),
                                     int(max_iters), int(error_type), int(bool(sym_check)), ctypes.c_double(laf_coef),
                                     int(bool(degen_check)), ctypes.c_uint64(int(seed)), _dp(F),
                                     mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                     stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if rc != 0:
        raise ValueError("port rejected the input (rc=%d)" % rc)
    return F.reshape(3, 3), mask.astype(bool), stats


def find_homography_raw(pts1, pts2, px_th=1.0, conf=0.999, max_iters=50000, error_type=0, sym_check=True, laf_coef=0.0,
                        seed=0, final_lsq=False):
    lib().port_set_final_lsq(int(bool(final_lsq)))
    p1 = np.ascontiguousarray(pts1, dtype=np.float64); p2 = np.ascontiguousarray(pts2, dtype=np.float64)
    n, dim = p1.shape
    H = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); stats = np.zeros(4, dtype=np.int32)
    rc = lib().port_find_homography(_dp(p1), _dp(p2), n, dim, ctypes.c_double(px_th), ctypes.c_double(conf),
                                    int(max_iters), int(error_type), int(bool(sym_check)), ctypes.c_double(laf_coef),
                                    ctypes.c_uint64(int(seed)), _dp(H),
                                    mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                    stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if rc != 0:
        raise ValueError("port rejected the input (rc=%d)" % rc)
    return H.reshape(3, 3), mask.astype(bool), stats
