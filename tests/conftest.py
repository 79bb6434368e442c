import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def norm_model(M):
    """Scale/sign normalisation used for every model comparison (F and H are defined up to scale, LAPACK/Jacobi
    eigenvector signs are arbitrary): unit Frobenius norm, largest-magnitude entry positive."""
    import numpy as np
    M = np.asarray(M, dtype=np.float64)
    n = np.linalg.norm(M)
    if n == 0:
        return M
    M = M / n
    return M * np.sign(M.flat[np.argmax(np.abs(M))])


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libdegensac_ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    return ref
