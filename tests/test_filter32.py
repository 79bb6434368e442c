"""The wave's FP32 upper-bound MSAC score (csrc/filter32.h) must never under-estimate the exact FP64 score and
must not change any result.  Checked through the one-thread host emulation, which is compiled with
DG_FILTER_CHECK: every model scored by the filter is also scored in FP64 and compared."""
import ctypes

import numpy as np
import pytest

from pydegensac_b200.scenes import scene_F


def _stats(E):
    ck = ctypes.c_long(); vi = ctypes.c_long(); ms = ctypes.c_double()
    E.emu_filter_stats(ctypes.byref(ck), ctypes.byref(vi), ctypes.byref(ms))
    return ck.value, vi.value, ms.value


@pytest.mark.parametrize("scale", [1.0, 8.0, 0.05])
def test_upper_bound_holds_and_results_unchanged(scale):
    from tests.hostemu import emu
    E = emu.lib()
    c0, v0, _ = _stats(E)
    for sc in range(4):
        for plane in (0.0, 0.8):
            for et in (0, 1):
                p1, p2, _ = scene_F(1500, 0.3, 100 + sc, plane)
                p1 = p1 * scale + 1000.0 * (scale - 1.0)      # other coordinate scales and large offsets
                p2 = p2 * scale - 300.0 * (scale - 1.0)
                kw = dict(px_th=1.0 * scale, conf=0.9999, max_iters=3000, error_type=et, seed=sc)
                E.emu_set_filter32(1); a = emu.find_fundamental(p1, p2, **kw)
                E.emu_set_filter32(0); b = emu.find_fundamental(p1, p2, **kw)
                E.emu_set_filter32(1)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    c1, v1, slack = _stats(E)
    assert c1 - c0 > 2000, "filter was not exercised"
    assert v1 - v0 == 0, "FP32 bound fell below the FP64 score %d times" % (v1 - v0)


@pytest.mark.parametrize("scale", [1.0, 6.0, 0.1])
def test_homography_upper_bound_holds_and_results_unchanged(scale):
    """Same contract for the H wave's FP32 filter (Sampson metric): J_up >= J on every scored model, identical outputs."""
    from tests.hostemu import emu
    from pydegensac_b200.scenes import scene_H
    E = emu.lib()
    ck = ctypes.c_long(); vi = ctypes.c_long()
    E.emu_hfilter_stats(ctypes.byref(ck), ctypes.byref(vi)); c0, v0 = ck.value, vi.value
    for sc in range(5):
        p1, p2, _ = scene_H(1200, 130 + 40 * sc, 20 + sc)
        p1 = p1 * scale + 2000.0 * (scale - 1.0)
        p2 = p2 * scale - 500.0 * (scale - 1.0)
        kw = dict(px_th=3.0 * scale, conf=0.9999, max_iters=6000, seed=sc)
        E.emu_set_filter32(1); a = emu.find_homography_raw(p1, p2, **kw)
        E.emu_set_filter32(0); b = emu.find_homography_raw(p1, p2, **kw)
        E.emu_set_filter32(1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    E.emu_hfilter_stats(ctypes.byref(ck), ctypes.byref(vi))
    assert ck.value - c0 > 150, "filter was not exercised"
    assert vi.value - v0 == 0, "FP32 bound fell below the FP64 score %d times" % (vi.value - v0)


@pytest.mark.parametrize("scale", [1.0, 5.0, 0.2])
def test_plane_and_parallax_count_bound_holds_and_results_unchanged(scale):
    """DEGENSAC's plane-and-parallax search (rFtH) counts the support of thousands of two-point hypotheses; only counts
    above the best so far matter.  The FP32 bound must never fall below the exact count, must settle nearly all of the
    hypotheses, and switching it off must not change a byte of the result."""
    from tests.hostemu import emu
    from pydegensac_b200.scenes import scene_F
    E = emu.lib()
    ck = ctypes.c_long(); vi = ctypes.c_long(); se = ctypes.c_long()
    E.emu_pp_stats(ctypes.byref(ck), ctypes.byref(vi), ctypes.byref(se)); c0, v0, s0 = ck.value, vi.value, se.value
    for sc in range(3):
        p1, p2, _ = scene_F(1500, 0.35, 40 + sc, 0.8)
        p1 = p1 * scale + 1000.0 * (scale - 1.0)
        p2 = p2 * scale - 300.0 * (scale - 1.0)
        kw = dict(px_th=1.0 * scale, conf=0.9999, max_iters=4000, seed=sc)
        E.emu_set_filter32(1); a = emu.find_fundamental(p1, p2, **kw)
        E.emu_set_filter32(0); b = emu.find_fundamental(p1, p2, **kw)
        E.emu_set_filter32(1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    E.emu_pp_stats(ctypes.byref(ck), ctypes.byref(vi), ctypes.byref(se))
    checked, settled = ck.value - c0, se.value - s0
    assert checked > 1000, "the plane-and-parallax search was not exercised"
    assert vi.value - v0 == 0, "FP32 count bound fell below the exact count %d times" % (vi.value - v0)
    assert settled > 0.9 * checked, "the bound settles too few hypotheses (%d of %d)" % (settled, checked)


def test_homography_residual_row_classification_is_safe():
    """The H driver builds its residual rows in two steps (Sampson metric): an FP32 lower bound of the error marks the
    correspondences that are at least 18 th away (they get +inf: no consumer of a row looks beyond 18 th), the rest is
    evaluated exactly.  Every correspondence marked 'far' must really be that far, and most outliers must be caught."""
    from tests.hostemu import emu
    from pydegensac_b200.scenes import scene_H
    E = emu.lib()
    pt = ctypes.c_long(); fa = ctypes.c_long(); vi = ctypes.c_long()
    E.emu_hres_stats(ctypes.byref(pt), ctypes.byref(fa), ctypes.byref(vi)); p0, f0, v0 = pt.value, fa.value, vi.value
    for sc, scale in enumerate((1.0, 7.0, 0.15)):
        p1, p2, _ = scene_H(1500, 300 + 100 * sc, 90 + sc)
        p1 = p1 * scale + 3000.0 * (scale - 1.0)
        p2 = p2 * scale - 800.0 * (scale - 1.0)
        kw = dict(px_th=3.0 * scale, conf=0.999, max_iters=3000, seed=sc)
        E.emu_set_filter32(1); a = emu.find_homography_raw(p1, p2, **kw)
        E.emu_set_filter32(0); b = emu.find_homography_raw(p1, p2, **kw)
        E.emu_set_filter32(1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    E.emu_hres_stats(ctypes.byref(pt), ctypes.byref(fa), ctypes.byref(vi))
    assert pt.value - p0 > 100000, "the classification was not exercised"
    assert vi.value - v0 == 0, "a correspondence marked far was closer than 18 th (%d times)" % (vi.value - v0)
    assert fa.value - f0 > 0.5 * (pt.value - p0)
