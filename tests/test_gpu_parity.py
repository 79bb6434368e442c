"""CUDA engine through the C ABI vs the compiled reference on the same seeded inputs (-m gpu), plus
size-independent properties at the benchmark's full size."""
import numpy as np
import pytest

from tests.conftest import norm_model
from pydegensac_b200.scenes import scene_F, scene_H, batch_F

pytestmark = pytest.mark.gpu


def _cmp(a, b, what):
    if np.abs(a[0]).sum() == 0 and np.abs(b[0]).sum() == 0:
        return
    assert np.array_equal(a[1], b[1]), "mask differs: " + what
    assert np.linalg.norm(norm_model(a[0]) - norm_model(b[0])) < 1e-6, what
    assert a[2][0] == b[2][0] and a[2][1] == b[2][1], what


def test_F_configs_vs_reference(ref_oracle):
    from pydegensac_b200 import _cabi
    for plane in (0.0, 0.8):
        p1, p2, _ = scene_F(2000, 0.3, 0, plane)
        for degen in (False, True):
            for seed in range(3):
                a = ref_oracle.find_fundamental(p1, p2, 1.0, 0.9999, 10000, degen_check=degen, seed=seed)
                F, m, s = _cabi.fundamental_batch(p1, p2, 1.0, 0.9999, 10000, 0, True, 0.0, degen, [seed])
                _cmp(a, (F[0], m[0], s[0]), "F plane=%s degen=%s seed=%d" % (plane, degen, seed))


def test_H_configs_vs_reference(ref_oracle):
    from pydegensac_b200 import _cabi
    p1, p2, _ = scene_H()
    for et in range(5):
        for seed in range(2):
            a = ref_oracle.find_homography_raw(p1, p2, 3.0, 0.999, 10000, error_type=et, seed=seed)
            H, m, s = _cabi.homography_batch(p1, p2, 3.0, 0.999, 10000, et, True, 0.0, [seed])
            _cmp(a, (H[0], m[0], s[0]), "H metric=%d seed=%d" % (et, seed))


def test_batch_of_distinct_pairs_vs_reference(ref_oracle):
    """Config 5 slice: a batch of distinct scenes, per-pair seeds; every pair must equal its single-pair reference run."""
    from pydegensac_b200 import _cabi
    P = 48
    b1, b2 = batch_F(P, 2000, 0.3, seed0=500)
    seeds = np.arange(P, dtype=np.uint64) + 500
    F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
    same = 0
    for i in range(P):
        a = ref_oracle.find_fundamental(b1[i], b2[i], 1.0, 0.9999, 10000, seed=int(seeds[i]))
        if np.array_equal(a[1], m[i]) and np.linalg.norm(norm_model(a[0]) - norm_model(F[i])) < 1e-6:
            same += 1
    # Bit-level agreement of every pair is not attainable: the 7-point cubic is solved through libm pow/acos/cos
    # (Ftools.c:272-294) whose last-bit differences between glibc and the CUDA math library are amplified by
    # ill-conditioned samples and the discontinuous DEGENSAC test.  Expect >= 95 %.
    print("identical to the reference: %d of %d pairs" % (same, P))
    assert same >= int(0.95 * P), "%d of %d pairs identical to the reference" % (same, P)


def test_randomised_small_configs_vs_reference(ref_oracle):
    from pydegensac_b200 import _cabi
    rng = np.random.default_rng(77)
    for case in range(40):
        kind = rng.choice(["F", "H"])
        n = int(rng.choice([8, 9, 12, 20, 50, 100, 300, 1000]))
        ratio = float(rng.choice([0.3, 0.5, 0.8, 1.0]))
        px = float(rng.choice([0.5, 1.0, 3.0])); conf = float(rng.choice([0.9, 0.99, 0.9999]))
        mi = int(rng.choice([50, 51, 100, 1000, 3000]))
        sym = bool(rng.integers(2)); seed = int(rng.integers(1 << 30)); sc = int(rng.integers(1000))
        if kind == "F":
            plane = float(rng.choice([0, 0, 0.5, 0.9])); et = int(rng.integers(2)); dg = bool(rng.integers(2))
            p1, p2, _ = scene_F(n, ratio, sc, plane)
            a = ref_oracle.find_fundamental(p1, p2, px, conf, mi, error_type=et, sym_check=sym, degen_check=dg, seed=seed)
            M, m, s = _cabi.fundamental_batch(p1, p2, px, conf, mi, et, sym, 0.0, dg, [seed])
        else:
            et = int(rng.integers(5))
            p1, p2, _ = scene_H(n, int(n * ratio), sc)
            a = ref_oracle.find_homography_raw(p1, p2, px, conf, mi, error_type=et, sym_check=sym, seed=seed)
            if a[2][3] <= 4 or a[2][2] >= a[2][0]:   # no consensus / every sample rejected: reference runs on uninitialised memory
                continue
            M, m, s = _cabi.homography_batch(p1, p2, px, conf, mi, et, sym, 0.0, [seed])
        _cmp(a, (M[0], m[0], s[0]), "case %d %s n=%d" % (case, kind, n))


def test_full_size_properties():
    """Size-independent properties on a benchmark-size batch: determinism (same seeds -> identical bytes),
    order independence (a pair's answer does not depend on its batch position), mask consistent with the
    returned model, rank-2 F, recall of the planted inliers."""
    from pydegensac_b200 import _cabi
    P = 256
    b1, b2 = batch_F(P, 2000, 0.3, seed0=0)
    seeds = np.arange(P, dtype=np.uint64)
    F, m, s = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
    F2, m2, s2 = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
    assert np.array_equal(F, F2) and np.array_equal(m, m2) and np.array_equal(s, s2)
    perm = np.random.default_rng(1).permutation(P)
    F3, m3, s3 = _cabi.fundamental_batch(b1[perm], b2[perm], 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds[perm])
    assert np.array_equal(F3, F[perm]) and np.array_equal(m3, m[perm])
    assert (s[:, 0] == 10000).all()           # config 2 never terminates early
    assert m[:, :600].mean() > 0.95           # planted inliers recovered
    assert m[:, 600:].mean() < 0.02
    for i in range(0, P, 16):
        sv = np.linalg.svd(F[i], compute_uv=False)
        assert sv[2] < 1e-10 * sv[0]          # rank 2
        x1 = np.concatenate([b1[i], np.ones((2000, 1))], 1); x2 = np.concatenate([b2[i], np.ones((2000, 1))], 1)
        l2 = x1 @ F[i].T; l1 = x2 @ F[i]
        r = np.sum(x2 * l2, 1)
        e = r * r / (l2[:, 0] ** 2 + l2[:, 1] ** 2 + l1[:, 0] ** 2 + l1[:, 1] ** 2)
        # mask == (Sampson <= th) up to the symmetric prune, which only clears entries (and, by the reference's
        # indexing quirk, clears list positions, i.e. low indices)
        assert not np.any(m[i] & (e > 1.0 + 1e-9))
        assert (m[i] != (e <= 1.0)).sum() <= 60


def test_python_api_on_gpu():
    import pydegensac_b200 as pdg
    p1, p2, gt = scene_H(1000, 400, 3)
    H, mask = pdg.findHomography(p1, p2, 3.0, seed=5)
    assert mask.dtype == bool and mask.sum() >= 390
    from pydegensac_b200.scenes import H_GT
    assert np.abs(H / H[2, 2] - H_GT).max() < 0.5
    q1, q2, gt = scene_F(1000, 0.5, 4)
    F, fm = pdg.findFundamentalMatrix(q1, q2, 1.0, 0.999, 5000, seed=9)
    assert fm[:500].mean() > 0.9
    F2, fm2 = pdg.findFundamentalMatrix(q1, q2, 1.0, 0.999, 5000, seed=9)
    assert np.array_equal(F, F2) and np.array_equal(fm, fm2)
    Fb, mb = pdg.findFundamentalMatrixBatch(np.stack([q1, q1]), np.stack([q2, q2]), 1.0, 0.999, 5000, seeds=[9, 9])
    assert np.array_equal(Fb[0], F) and np.array_equal(mb[1], np.asarray(fm))
    Hb, hmb = pdg.findHomographyBatch(p1[None], p2[None], 3.0, seeds=[5])
    assert np.allclose(Hb[0], H) and np.array_equal(hmb[0], mask)
    # degenerate input: all points identical -> no model, all-False mask, no exception
    z = np.ones((20, 2))
    Fz, mz = pdg.findFundamentalMatrix(z, z, 1.0, 0.99, 100, seed=1)
    assert np.abs(Fz).sum() == 0 and not any(mz)
    # [N,6] keypoints with shapes + laf_consistensy_coef through the public API (reference signature, utils.py:111)
    import warnings
    from pydegensac_b200 import _cabi
    from pydegensac_b200.scenes import scene_F_laf
    x1, x2, _ = scene_F_laf(400, 0.6, 2)
    Fl, ml = pdg.findFundamentalMatrix(x1, x2, 1.0, 0.999, 2000, laf_consistensy_coef=2.0, seed=3)
    Fc, mc, _ = _cabi.fundamental_batch(x1, x2, 1.0, 0.999, 2000, 0, True, 2.0, True, [3])
    assert np.array_equal(Fl, Fc[0]) and np.array_equal(np.asarray(ml), mc[0])
    with warnings.catch_warnings(record=True) as wlist:    # (x,y) only: the coefficient is dropped with a warning
        warnings.simplefilter("always")
        F0, m0 = pdg.findFundamentalMatrix(x1[:, :2], x2[:, :2], 1.0, 0.999, 2000, laf_consistensy_coef=2.0, seed=3)
        assert any("laf" in str(w.message).lower() for w in wlist)
    F1, m1 = pdg.findFundamentalMatrix(x1[:, :2], x2[:, :2], 1.0, 0.999, 2000, seed=3)
    assert np.array_equal(F0, F1) and np.array_equal(np.asarray(m0), np.asarray(m1))
    # batched entry points take the coefficient too
    Fb2, mb2 = pdg.findFundamentalMatrixBatch(np.stack([x1, x1]), np.stack([x2, x2]), 1.0, 0.999, 2000, seeds=[3, 3],
                                              laf_consistensy_coef=2.0)
    assert np.array_equal(Fb2[0], Fl) and np.array_equal(mb2[1], np.asarray(ml))


def test_host_feed_fallback_gives_identical_results():
    """The host-buffer API streams its input while the kernel runs; CTAs that outwait their patience retire and a
    second launch finishes the batch (what happens when a profiler serialises the streams).  Forcing that path
    (DGB200_FEED_WAIT_US=0, read per call) must not change a single output byte.  The batch is large enough for the
    chunked feed (several chunks beyond the first 592 pairs)."""
    import os
    from pydegensac_b200 import _cabi
    P = 1500
    b1, b2 = batch_F(8, 600, 0.5, seed0=900)
    b1 = np.ascontiguousarray(np.tile(b1, (P // 8 + 1, 1, 1))[:P]); b2 = np.ascontiguousarray(np.tile(b2, (P // 8 + 1, 1, 1))[:P])
    seeds = np.arange(P, dtype=np.uint64) % 8
    a = _cabi.fundamental_batch(b1, b2, 1.0, 0.99, 300, 0, True, 0.0, True, seeds)
    os.environ["DGB200_FEED_WAIT_US"] = "0"
    try:
        b = _cabi.fundamental_batch(b1, b2, 1.0, 0.99, 300, 0, True, 0.0, True, seeds)
    finally:
        del os.environ["DGB200_FEED_WAIT_US"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # identical scenes with identical seeds give identical rows (batch-order independence of the persistent CTAs)
    for i in range(8, P):
        assert np.array_equal(a[1][i], a[1][i % 8])


def test_F_laf_gate_vs_reference_on_gpu(ref_oracle):
    """LAF-consistency gate through the C ABI ([N,6] inputs, laf_coef > 0) against the compiled reference."""
    from pydegensac_b200 import _cabi
    from pydegensac_b200.scenes import scene_F_laf
    rng = np.random.default_rng(5)
    checked = 0
    for case in range(36):
        n = int(rng.choice([100, 300, 600, 1000])); ratio = float(rng.choice([0.4, 0.6, 0.8])); seed = int(rng.integers(1 << 20))
        jitter = float(rng.choice([0.2, 0.6, 1.5])); laf = float(rng.choice([0.5, 1.0, 2.0, 5.0])); et = int(rng.integers(2))
        sym = bool(rng.integers(2)); mi = int(rng.choice([200, 1000, 3000])); plane = float(rng.choice([0, 0, 0.6]))
        x1, x2, _ = scene_F_laf(n, ratio, seed, jitter, plane)
        a = ref_oracle.find_fundamental(x1, x2, 1.0, 0.999, mi, error_type=et, sym_check=sym, laf_coef=laf, degen_check=True, seed=seed)
        if a[2][3] <= 4 or a[2][2] >= a[2][0]:
            continue
        F, m, s = _cabi.fundamental_batch(x1, x2, 1.0, 0.999, mi, et, sym, laf, True, [seed])
        _cmp(a, (F[0], m[0], s[0]), "F LAF case %d" % case)
        checked += 1
    assert checked >= 25
    # [N,2] inputs with laf_coef > 0 are rejected at the C ABI (the Python layer warns and drops the coefficient)
    with pytest.raises(Exception):
        _cabi.fundamental_batch(x1[:, :2].copy(), x2[:, :2].copy(), 1.0, 0.999, 100, 0, True, 1.0, True, [1])


def test_H_laf_gate_vs_reference_on_gpu(ref_oracle):
    """LAF gate of the homography path through the C ABI, all five metrics."""
    from pydegensac_b200 import _cabi
    from pydegensac_b200.scenes import scene_H_laf
    rng = np.random.default_rng(9)
    for case in range(40):
        n = int(rng.choice([100, 300, 800])); nin = int(n * float(rng.choice([0.4, 0.6, 0.8]))); seed = int(rng.integers(1 << 20))
        jitter = float(rng.choice([0.2, 0.6, 1.5])); laf = float(rng.choice([0.5, 1.0, 2.0, 5.0, 20.0])); et = int(rng.integers(5))
        sym = bool(rng.integers(2)); mi = int(rng.choice([200, 1000, 3000])); px = float(rng.choice([1.0, 3.0]))
        x1, x2, _ = scene_H_laf(n, nin, seed, jitter)
        a = ref_oracle.find_homography_raw(x1, x2, px, 0.999, mi, error_type=et, sym_check=sym, laf_coef=laf, seed=seed)
        H, m, s = _cabi.homography_batch(x1, x2, px, 0.999, mi, et, sym, laf, [seed])
        _cmp(a, (H[0], m[0], s[0]), "H LAF case %d" % case)
