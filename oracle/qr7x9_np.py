"""Oracle of the reference's alternative 7-point null-space solver nullspace_qr7x9 (Ftools.c:594-668, USE_QR) --
TEST INFRASTRUCTURE ONLY.

The reference calls LAPACK dgeqp3_ with `lapack_int = ptrdiff_t` pivot arrays (lapwrap.h:12); against the LP64 LAPACK
available in this image that call corrupts the pivot vector, so the compiled leaf cannot be run
here.  This restatement uses the same LAPACK routine through scipy.linalg.qr(pivoting=True) (dgeqp3) and then follows
the reference's back substitution literally (Ftools.c:640-666)."""
import numpy as np


def nullspace_qr7x9(A):
    from scipy.linalg import qr
    A = np.asarray(A, dtype=np.float64).reshape(7, 9)
    R, p = qr(A, mode="r", pivoting=True)          # R: 7 x 9 upper trapezoidal of A[:, p]
    rows, cols = 7, 9
    N = np.zeros((2, 9))
    for k in (1, 2):
        sol = N[k - 1]
        for c in range(rows, cols):
            sol[p[c]] = 0.0
        sol[p[cols - k]] = 1.0
        for r in range(rows - 1, -1, -1):
            if R[r, r] == 0.0:
                return -1, N
            a = 0.0
            for c in range(r + 1, cols):
                a += R[r, c] * sol[p[c]]
            sol[p[r]] = -a / R[r, r]
    return 0, N
