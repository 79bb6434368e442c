"""Small linear algebra that replaces LAPACK / CCMATH on the path (host build of the engine headers):
the SVD-free rank-2 projection (dominant eigenvector of adj(F)^T adj(F) by repeated squaring) against numpy's SVD and
against the Jacobi route it falls back to."""
import ctypes

import numpy as np
import pytest

from tests.hostemu import emu


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _rank2_numpy(F):
    U, s, Vt = np.linalg.svd(F)
    s[2] = 0.0
    return U @ np.diag(s) @ Vt


@pytest.mark.parametrize("kind", ["random", "unnormalised", "near_rank2", "clustered"])
def test_rank2_projection_matches_svd(kind):
    L = emu.lib()
    rng = np.random.default_rng(11)
    worst = 0.0
    fast = 0
    for t in range(400):
        F = rng.uniform(-0.5, 0.5, (3, 3))
        if kind == "unnormalised":          # entries spanning 12 orders of magnitude, as an un-normalised 8-point F
            sc = np.array([1e-6, 1e-3, 1.0])
            F = F * sc[:, None] * sc[None, :]
        elif kind == "near_rank2":
            U, s, Vt = np.linalg.svd(F)
            s[2] *= 1e-7
            F = U @ np.diag(s) @ Vt
        elif kind == "clustered":           # two smallest singular values within 5 %: the fast path must decline
            U, s, Vt = np.linalg.svd(F)
            s[2] = 0.97 * s[1]
            F = U @ np.diag(s) @ Vt
        F = np.ascontiguousarray(F)
        v = np.zeros(3)
        fast += L.emu_smallest_right_sv3_fast(_dp(F.copy()), _dp(v))
        A = F.copy()
        L.emu_enforce_rank2(_dp(A))
        B = F.copy()
        L.emu_enforce_rank2_slow(_dp(B))
        ref = _rank2_numpy(F)
        n = np.linalg.norm(ref)
        worst = max(worst, np.linalg.norm(A - ref) / n, np.linalg.norm(B - ref) / n)
        assert np.linalg.matrix_rank(A, tol=1e-9 * n) <= 2
    # clustered case: the projection itself is ill-conditioned (eps / relative gap), everything else is ~1e-14
    assert worst < (2e-10 if kind == "clustered" else 5e-12), worst
    if kind == "clustered":
        assert fast < 400          # some of them must have taken the Jacobi fallback
    else:
        assert fast > 380          # the SVD-free path handles (nearly) all ordinary matrices
