"""pydegensac_b200 -- B200-native LO-RANSAC / DEGENSAC behind the pydegensac API.

    import pydegensac_b200 as pydegensac
    H, mask = pydegensac.findHomography(src_pts, dst_pts, 4.0, 0.99, 2000)
    F, mask = pydegensac.findFundamentalMatrix(src_pts, dst_pts, 0.5, 0.999, 50000)

Mirrors src/pydegensac/__init__.py:1-4 of the reference: the native `findHomography_` /
`findFundamentalMatrix_` (pybind11 over the C ABI in include/degensac_b200.h) are re-exported when the
native build is present; the compute path is CUDA-only and raises if it is not.
"""
from .utils import (findHomography,
                    findFundamentalMatrix,
                    findHomographyBatch,
                    findFundamentalMatrixBatch,
                    findHomographyFromEllipses,
                    laf_to_ellipse_frame,
                    convert_cv2_kpts_to_xyA,
                    error_type_dict_homography,
                    error_type_dict_fundamental)

try:  # `from .pydegensac import *` of the reference; absent until `python -m pydegensac_b200.build` has run
    from .pydegensac import findHomography_, findFundamentalMatrix_  # noqa: F401
except ImportError:  # pragma: no cover
    def _missing(*a, **k):
        raise RuntimeError("pydegensac_b200: native module not built - run `python -m pydegensac_b200.build` "
                           "(CUDA-only engine, no CPU fallback)")
    findHomography_ = findFundamentalMatrix_ = _missing

__version__ = "0.1.0"
