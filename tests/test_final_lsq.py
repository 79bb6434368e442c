"""DGB200_FLAG_FINAL_LSQ = the reference's compile-time __FINAL_LSQ__ (exp_ranF.h:28-29, exp_ranF.c:1701-1705,
exp_ranH.c:866-870).

Oracles: for H the UNMODIFIED reference compiled with -D__FINAL_LSQ__ (oracle/_ref/libdegensac_ref_lsq.so); for F the
reference text under that macro does not compile (a Score assigned to an unsigned, exp_ranF.c:1702), so the plain-C
restatement oracle/port -- pinned to the reference on the H side by the first test -- carries it."""
import numpy as np
import pytest

from tests.conftest import norm_model
from pydegensac_b200.scenes import scene_F, scene_H


def _same(a, b, what, tol=1e-6):
    assert np.array_equal(a[1], b[1]), "mask differs: " + what
    assert np.linalg.norm(norm_model(a[0]) - norm_model(np.asarray(b[0]).reshape(3, 3))) < tol, what


def test_port_final_lsq_matches_reference_build_H(ref_oracle):
    from oracle import port
    if not ref_oracle.available_final_lsq():
        pytest.skip("libdegensac_ref_lsq.so not built")
    changed = 0
    for sc in range(6):
        for et in (0, 1, 3):
            p1, p2, _ = scene_H(700, 300, sc)
            a = ref_oracle.find_homography_raw(p1, p2, 3.0, 0.999, 1500, error_type=et, seed=sc, final_lsq=True)
            b = port.find_homography_raw(p1, p2, 3.0, 0.999, 1500, error_type=et, seed=sc, final_lsq=True)
            _same(a, b, "H scene %d metric %d" % (sc, et))
            plain = ref_oracle.find_homography_raw(p1, p2, 3.0, 0.999, 1500, error_type=et, seed=sc)
            changed += int(not np.array_equal(plain[0], a[0]))
    assert changed > 0, "the polish never changed a model: option not exercised"


def test_host_emulation_final_lsq_vs_oracles(ref_oracle):
    from oracle import port
    from tests.hostemu import emu
    E = emu.lib()
    try:
        E.emu_set_final_lsq(1)
        for sc in range(5):
            p1, p2, _ = scene_F(900, 0.4, sc, 0.5 if sc % 2 else 0.0)
            for et in (0, 1):
                a = port.find_fundamental(p1, p2, 1.0, 0.999, 1500, error_type=et, seed=sc, final_lsq=True)
                b = emu.find_fundamental(p1, p2, 1.0, 0.999, 1500, error_type=et, seed=sc)
                _same(a, b, "F scene %d metric %d" % (sc, et))
        if ref_oracle.available_final_lsq():
            for sc in range(4):
                p1, p2, _ = scene_H(800, 350, sc)
                a = ref_oracle.find_homography_raw(p1, p2, 3.0, 0.999, 1500, seed=sc, final_lsq=True)
                b = emu.find_homography_raw(p1, p2, 3.0, 0.999, 1500, seed=sc)
                _same(a, b, "H scene %d" % sc)
    finally:
        E.emu_set_final_lsq(0)


@pytest.mark.gpu
def test_gpu_final_lsq_vs_oracles(ref_oracle):
    import pydegensac_b200 as pdg
    from pydegensac_b200 import _cabi
    from oracle import port
    P = 6
    for et in (0, 1):
        scenes = [scene_F(2000, 0.3, 40 + i, 0.8 if i % 3 == 0 else 0.0) for i in range(P)]
        b1 = np.stack([s[0] for s in scenes]); b2 = np.stack([s[1] for s in scenes])
        seeds = np.arange(P, dtype=np.uint64) + 40
        F, m, st = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 4000, et, True, 0.0, True, seeds, flags=_cabi.FLAG_FINAL_LSQ)
        F0, m0, _ = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 4000, et, True, 0.0, True, seeds)
        assert not np.array_equal(F, F0)
        for i in range(P):
            a = port.find_fundamental(b1[i], b2[i], 1.0, 0.9999, 4000, error_type=et, seed=int(seeds[i]), final_lsq=True)
            _same(a, (F[i], m[i]), "gpu F pair %d metric %d" % (i, et))
    if ref_oracle.available_final_lsq():
        for et in range(5):
            p1, p2, _ = scene_H(3000, 1000, et)
            a = ref_oracle.find_homography_raw(p1, p2, 3.0, 0.999, 4000, error_type=et, seed=et, final_lsq=True)
            H, m, st = _cabi.homography_batch(p1, p2, 3.0, 0.999, 4000, et, True, 0.0, [et], flags=_cabi.FLAG_FINAL_LSQ)
            _same(a, (H[0], m[0]), "gpu H metric %d" % et)
    # public API keyword
    p1, p2, _ = scene_F(1000, 0.5, 3)
    Fp, mp = pdg.findFundamentalMatrixBatch(p1[None], p2[None], 1.0, 0.999, 2000, seeds=[3], final_lsq=True)
    a = port.find_fundamental(p1, p2, 1.0, 0.999, 2000, seed=3, final_lsq=True)
    _same(a, (Fp[0], mp[0]), "public API")
