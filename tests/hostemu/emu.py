"""ctypes loader of the one-thread host emulation of the engine (TEST HELPER ONLY)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostemu.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "hostemu.cpp")
    csrc = os.path.join(_HERE, "..", "..", "pydegensac_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith(".h")])
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", _SO, src])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def find_fundamental(p1, p2, px_th=0.5, conf=0.9999, max_iters=100000, error_type=0, sym_check=True, laf_coef=0.0,
                     degen_check=True, seed=0, chunk=512):
    p1 = np.ascontiguousarray(p1, dtype=np.float64); p2 = np.ascontiguousarray(p2, dtype=np.float64)
    n, dim = p1.shape
    F = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_find_fundamental(_dp(p1), _dp(p2), n, dim, ctypes.c_double(px_th), ctypes.c_double(conf),
                                    int(max_iters), int(error_type), int(bool(sym_check)), ctypes.c_double(laf_coef),
                                    int(bool(degen_check)), ctypes.c_uint64(seed), int(chunk), _dp(F),
                                    mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                    stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0, rc
    return F.reshape(3, 3), mask.astype(bool), stats


def find_homography_raw(p1, p2, px_th=1.0, conf=0.999, max_iters=50000, error_type=0, sym_check=True, laf_coef=0.0,
                        seed=0, chunk=512):
    p1 = np.ascontiguousarray(p1, dtype=np.float64); p2 = np.ascontiguousarray(p2, dtype=np.float64)
    n, dim = p1.shape
    H = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_find_homography(_dp(p1), _dp(p2), n, dim, ctypes.c_double(px_th), ctypes.c_double(conf),
                                   int(max_iters), int(error_type), int(bool(sym_check)), ctypes.c_double(laf_coef),
                                   ctypes.c_uint64(seed), int(chunk), _dp(H),
                                   mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                   stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0, rc
    return H.reshape(3, 3), mask.astype(bool), stats


def find_homography_2el_raw(u10, px_th=1.0, conf=0.999, max_iters=50000, seed=0, chunk=512):
    u = np.ascontiguousarray(u10, dtype=np.float64)
    n = u.shape[0]
    H = np.zeros(9); mask = np.zeros(n, dtype=np.uint8); stats = np.zeros(4, dtype=np.int32)
    rc = lib().emu_find_homography_2el(_dp(u), n, ctypes.c_double(px_th), ctypes.c_double(conf), int(max_iters),
                                       ctypes.c_uint64(seed), int(chunk), _dp(H),
                                       mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                       stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0, rc
    return H.reshape(3, 3), mask.astype(bool), stats


def h_from_2el(ua, ub):
    ua = np.ascontiguousarray(ua, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
    h = np.zeros(9)
    ok = lib().emu_h_from_2el(_dp(ua), _dp(ub), _dp(h))
    return bool(ok), h


def u2h4_calls(reset=True):
    """How often the fit took the reference's len == 4 branch of u2h (Htools.c:108-116), whose outcome depends on nine
    uninitialised stack doubles in the reference."""
    f = lib().emu_u2h4_calls
    f.restype = ctypes.c_long
    return int(f(int(reset)))
