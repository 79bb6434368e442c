"""Golden vectors (tests/golden/*.npz, produced by the UNMODIFIED reference through oracle/_ref with the
Philox replay stream) against: the reference itself (pins the fixtures), the one-thread host emulation of the
engine (CPU, always) and the CUDA engine through the C ABI (-m gpu)."""
import json
import os

import numpy as np
import pytest

from tests.conftest import norm_model
from pydegensac_b200.scenes import scene_F, scene_H

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
META = json.loads(str(G["meta"]))
TOL = 1e-6   # north_star: H/F within 1e-6 relative Frobenius after scale/sign normalisation


def _inputs(m):
    if m["kind"] == "F":
        p1, p2, _ = scene_F(**m["scene"])
    else:
        p1, p2, _ = scene_H(**m["scene"])
    return p1, p2


def _check(i, model, mask, stats):
    gm, gmask, gst = G["model_%d" % i], G["mask_%d" % i], G["stats_%d" % i]
    if np.abs(gm).sum() == 0:
        assert np.abs(model).sum() == 0
        return
    assert np.array_equal(np.asarray(mask, bool), gmask), "inlier mask differs from the reference (case %d)" % i
    assert np.linalg.norm(norm_model(model) - norm_model(gm)) < TOL
    assert int(stats[0]) == int(gst[0]) and int(stats[1]) == int(gst[1]), "samples drawn / LO runs differ"


@pytest.mark.parametrize("i", range(len(META)))
def test_reference_reproduces_golden(i, ref_oracle):
    m = META[i]
    p1, p2 = _inputs(m)
    if m["kind"] == "F":
        M, mask, st = ref_oracle.find_fundamental(p1, p2, **m["call"])
    else:
        M, mask, st = ref_oracle.find_homography_raw(p1, p2, **m["call"])
    assert np.array_equal(mask, G["mask_%d" % i])
    assert np.allclose(M, G["model_%d" % i], rtol=0, atol=0)


@pytest.mark.parametrize("i", range(len(META)))
def test_host_emulation_matches_golden(i):
    from tests.hostemu import emu
    m = META[i]
    p1, p2 = _inputs(m)
    kw = dict(m["call"])
    if m["kind"] == "F":
        M, mask, st = emu.find_fundamental(p1, p2, **kw)
    else:
        M, mask, st = emu.find_homography_raw(p1, p2, **kw)
    _check(i, M, mask, st)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META)))
def test_cuda_matches_golden(i):
    from pydegensac_b200 import _cabi
    m = META[i]
    p1, p2 = _inputs(m)
    c = m["call"]
    if m["kind"] == "F":
        M, mask, st = _cabi.fundamental_batch(p1, p2, c["px_th"], c["conf"], c["max_iters"], c["error_type"],
                                              c["sym_check"], 0.0, c["degen_check"], [c["seed"]])
    else:
        M, mask, st = _cabi.homography_batch(p1, p2, c["px_th"], c["conf"], c["max_iters"], c["error_type"],
                                             c["sym_check"], 0.0, [c["seed"]])
    _check(i, M[0], mask[0], st[0])


def test_dogman_host_emulation():
    """BASELINE.json configs[0]: the simple-example.py point set (frozen), H th 4 / conf .99 / 2000 iters."""
    from tests.hostemu import emu
    D = np.load(os.path.join(HERE, "golden", "dogman_v1.npz"))
    for seed in (0, 1):
        H, mask, st = emu.find_homography_raw(D["src"], D["dst"], 4.0, 0.99, 2000, seed=seed)
        assert np.array_equal(mask, D["H_mask_%d" % seed])
        assert np.linalg.norm(norm_model(H) - norm_model(D["H_raw_%d" % seed])) < TOL
        Hcv = np.linalg.inv(H.T)
        Hg = D["H_gt"] / D["H_gt"][2, 2]   # HPatches ground truth: agreement to ~1e-2
        assert np.linalg.norm(Hcv / Hcv[2, 2] - Hg) / np.linalg.norm(Hg) < 0.03


@pytest.mark.gpu
def test_dogman_cuda():
    import pydegensac_b200 as pdg
    D = np.load(os.path.join(HERE, "golden", "dogman_v1.npz"))
    for seed in (0, 1):
        H, mask = pdg.findHomography(D["src"], D["dst"], 4.0, 0.99, 2000, seed=seed)
        assert np.array_equal(np.asarray(mask), D["H_mask_%d" % seed])
        Href = np.linalg.inv(D["H_raw_%d" % seed].T)
        assert np.linalg.norm(norm_model(H) - norm_model(Href)) < TOL
        F, fmask = pdg.findFundamentalMatrix(D["src"], D["dst"], 0.5, 0.999, 50000, seed=seed)
        assert np.array_equal(np.asarray(fmask), D["F_mask_%d" % seed])
        assert np.linalg.norm(norm_model(F) - norm_model(D["F_%d" % seed])) < TOL


# ------------------------------------------------------------------ LAF-consistency gate ([N,6] inputs, laf_coef > 0)
GL = np.load(os.path.join(HERE, "golden", "golden_laf_v1.npz"))
META_L = json.loads(str(GL["meta"]))


def _inputs_laf(m):
    from pydegensac_b200.scenes import scene_F_laf, scene_H_laf
    return (scene_F_laf if m["kind"] == "F" else scene_H_laf)(**m["scene"])[:2]


def _check_laf(i, model, mask, stats):
    gm, gmask, gst = GL["model_%d" % i], GL["mask_%d" % i], GL["stats_%d" % i]
    assert np.array_equal(np.asarray(mask, bool), gmask.astype(bool)), "inlier mask differs from the reference (LAF case %d)" % i
    assert np.linalg.norm(norm_model(model) - norm_model(gm)) < TOL
    assert int(stats[0]) == int(gst[0]) and int(stats[1]) == int(gst[1]), "samples drawn / LO runs differ"


@pytest.mark.parametrize("i", range(len(META_L)))
def test_reference_reproduces_golden_laf(i, ref_oracle):
    m = META_L[i]
    p1, p2 = _inputs_laf(m)
    if m["kind"] == "F":
        out = ref_oracle.find_fundamental(p1, p2, **m["call"])
    else:
        out = ref_oracle.find_homography_raw(p1, p2, **m["call"])
    _check_laf(i, *out)


@pytest.mark.parametrize("i", range(len(META_L)))
def test_host_emulation_matches_golden_laf(i):
    from tests.hostemu import emu
    m = META_L[i]
    p1, p2 = _inputs_laf(m)
    k = m["call"]
    if m["kind"] == "F":
        out = emu.find_fundamental(p1, p2, k["px_th"], k["conf"], k["max_iters"], k["error_type"], k["sym_check"], k["laf_coef"],
                                   k["degen_check"], k["seed"])
    else:
        out = emu.find_homography_raw(p1, p2, k["px_th"], k["conf"], k["max_iters"], k["error_type"], k["sym_check"], k["laf_coef"],
                                      k["seed"])
    _check_laf(i, *out)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(META_L)))
def test_cuda_engine_matches_golden_laf(i):
    from pydegensac_b200 import _cabi
    m = META_L[i]
    p1, p2 = _inputs_laf(m)
    k = m["call"]
    if m["kind"] == "F":
        M, mask, st = _cabi.fundamental_batch(p1, p2, k["px_th"], k["conf"], k["max_iters"], k["error_type"], k["sym_check"],
                                              k["laf_coef"], k["degen_check"], [k["seed"]])
    else:
        M, mask, st = _cabi.homography_batch(p1, p2, k["px_th"], k["conf"], k["max_iters"], k["error_type"], k["sym_check"],
                                             k["laf_coef"], [k["seed"]])
    _check_laf(i, M[0], mask[0], st[0])
