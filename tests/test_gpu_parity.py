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
    bad = []
    for i in range(P):
        a = ref_oracle.find_fundamental(b1[i], b2[i], 1.0, 0.9999, 10000, seed=int(seeds[i]))
        if not (np.array_equal(a[1], m[i]) and np.linalg.norm(norm_model(a[0]) - norm_model(F[i])) < 1e-6
                and a[2][0] == s[i][0] and a[2][1] == s[i][1]):
            bad.append((i, int(a[1].sum()), int(m[i].sum())))
    # Every pair: identical mask, model within 1e-6, identical samples-drawn and LO-run counts (tools/parity_rate.py
    # runs the same check on 1024 + 256 pairs; its output is kept under profiles/).
    assert not bad, "pairs differing from the reference (index, ref inliers, gpu inliers): %s" % bad


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


def test_ragged_batch_equals_per_pair_calls(ref_oracle):
    """Ragged-N batch (offsets ABI): every pair equals its own single call and the reference."""
    from pydegensac_b200 import _cabi
    import pydegensac_b200 as pdg
    rng = np.random.default_rng(5)
    ns = [8, 37, 500, 2000, 1203, 64, 999, 2000, 15, 300]
    l1, l2 = [], []
    for i, n in enumerate(ns):
        p1, p2, _ = scene_F(n, 0.5, 100 + i)
        l1.append(p1); l2.append(p2)
    seeds = np.arange(len(ns), dtype=np.uint64) + 9
    F, masks, stats = _cabi.fundamental_ragged(l1, l2, 1.0, 0.999, 2000, 0, True, 0.0, True, seeds)
    for i, n in enumerate(ns):
        Fi, mi, si = _cabi.fundamental_batch(l1[i], l2[i], 1.0, 0.999, 2000, 0, True, 0.0, True, [int(seeds[i])])
        assert masks[i].shape == (n,)
        assert np.array_equal(Fi[0], F[i]) and np.array_equal(mi[0], masks[i]) and np.array_equal(si[0], stats[i]), "pair %d" % i
        if n >= 30:   # (tiny sets: the reference may run its LO on uninitialised memory)
            a = ref_oracle.find_fundamental(l1[i], l2[i], 1.0, 0.999, 2000, seed=int(seeds[i]))
            _cmp(a, (F[i], masks[i], stats[i]), "ragged F pair %d n=%d" % (i, n))
    # homographies, through the public Python entry point
    h1, h2 = [], []
    hs = [4, 50, 811, 5000, 129]
    for i, n in enumerate(hs):
        p1, p2, _ = scene_H(n, max(4, n // 2), 30 + i)
        h1.append(p1); h2.append(p2)
    H, hm = pdg.findHomographyBatch(h1, h2, 3.0, 0.999, 2000, seeds=np.arange(len(hs), dtype=np.uint64))
    for i, n in enumerate(hs):
        Hi, mi = pdg.findHomographyBatch(h1[i][None], h2[i][None], 3.0, 0.999, 2000, seeds=[i])
        assert np.array_equal(np.asarray(hm[i]), mi[0]) and np.allclose(H[i], Hi[0], rtol=0, atol=0), "H pair %d" % i


def test_device_entry_points_are_reentrant_across_streams():
    """F and H launches in flight on two streams at once give the same bytes as the same launches run one by one
    (every launch owns its scratch slabs and work counter)."""
    import torch
    from pydegensac_b200 import _cabi
    dev = torch.device("cuda:0")
    P, N = 96, 2000
    b1, b2 = batch_F(P, N, 0.3, seed0=40)
    d1 = torch.from_numpy(b1).to(dev); d2 = torch.from_numpy(b2).to(dev)
    q1, q2, _ = scene_H(N, 700, 2)
    e1 = torch.from_numpy(np.repeat(q1[None], P, 0).copy()).to(dev); e2 = torch.from_numpy(np.repeat(q2[None], P, 0).copy()).to(dev)
    seeds = torch.arange(P, dtype=torch.int64, device=dev)

    def outs():
        return (torch.zeros((P, 9), dtype=torch.float64, device=dev), torch.zeros((P, N), dtype=torch.uint8, device=dev),
                torch.zeros((P, 4), dtype=torch.int32, device=dev))

    def run(streamF, streamH):
        oF, oH = outs(), outs()
        for _ in range(2):   # two rounds back to back on each stream
            _cabi.fundamental_batch_dev(d1.data_ptr(), d2.data_ptr(), P, N, 2, 1.0, 0.9999, 3000, 0, True, 0.0, True,
                                        seeds.data_ptr(), oF[0].data_ptr(), oF[1].data_ptr(), oF[2].data_ptr(), streamF.cuda_stream)
            _cabi.homography_batch_dev(e1.data_ptr(), e2.data_ptr(), P, N, 2, 3.0, 0.999, 3000, 0, True, 0.0,
                                       seeds.data_ptr(), oH[0].data_ptr(), oH[1].data_ptr(), oH[2].data_ptr(), streamH.cuda_stream)
        torch.cuda.synchronize()
        return [t.cpu().numpy() for t in oF + oH]

    s0 = torch.cuda.current_stream()
    serial = run(s0, s0)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    overlapped = run(sa, sb)
    for x, y in zip(serial, overlapped):
        assert np.array_equal(x, y)


def test_fp32_wave_filter_does_not_change_results(monkeypatch):
    """DGB200_FILTER32=0 scores the hypothesis wave in FP64: outputs must be byte-identical (the FP32 score is an upper
    bound, so the filter only drops models the exact replay would have rejected)."""
    from pydegensac_b200 import _cabi
    P = 128
    b1, b2 = batch_F(P, 2000, 0.3, seed0=7)
    seeds = np.arange(P, dtype=np.uint64) + 7
    on = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
    monkeypatch.setenv("DGB200_FILTER32", "0")
    off = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
    monkeypatch.delenv("DGB200_FILTER32")
    for x, y in zip(on, off):
        assert np.array_equal(x, y)
    p1, p2, _ = scene_F(2000, 0.3, 3, 0.8)
    on = _cabi.fundamental_batch(p1, p2, 0.7, 0.9999, 10000, 1, True, 0.0, True, [3])   # symmetric epipolar metric
    monkeypatch.setenv("DGB200_FILTER32", "0")
    off = _cabi.fundamental_batch(p1, p2, 0.7, 0.9999, 10000, 1, True, 0.0, True, [3])
    for x, y in zip(on, off):
        assert np.array_equal(x, y)


def test_plane_and_parallax_count_bound_does_not_change_results(monkeypatch):
    """Dominant-plane scenes spend their time in DEGENSAC's plane-and-parallax search, whose two-point hypotheses are
    settled by an FP32 upper bound of their support on a packed tile of the off-plane correspondences (odd and even
    list lengths).  Without the filter tile (DGB200_FILTER32=0) every hypothesis is counted exactly in FP64: the outputs
    must be byte-identical."""
    from pydegensac_b200 import _cabi
    for (n, frac, plane, P) in [(2000, 0.3, 0.8, 48), (701, 0.4, 0.6, 32), (300, 0.5, 0.9, 32)]:
        b1, b2 = batch_F(P, n, frac, seed0=40, plane_frac=plane)
        seeds = np.arange(P, dtype=np.uint64) + 40
        monkeypatch.delenv("DGB200_FILTER32", raising=False)
        on = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
        monkeypatch.setenv("DGB200_FILTER32", "0")
        off = _cabi.fundamental_batch(b1, b2, 1.0, 0.9999, 10000, 0, True, 0.0, True, seeds)
        monkeypatch.delenv("DGB200_FILTER32")
        for x, y in zip(on, off):
            assert np.array_equal(x, y)
        assert (on[2][:, 2] > 0).sum() >= P // 4, "the plane branch was hardly exercised"   # stats[2] = plane inliers found


def test_fp32_wave_filter_homography(monkeypatch, ref_oracle):
    """The H wave's FP32 filter (Sampson metric): byte-identical outputs with the filter on and off at N = 5000 (tile in
    the slab) and N = 1200 (tile in shared memory), odd N included (padding slot of the pair-interleaved tile)."""
    from pydegensac_b200 import _cabi
    for (n, n_in, P) in [(5000, 1500, 24), (1201, 300, 32), (4097, 900, 8)]:
        b1 = np.empty((P, n, 2)); b2 = np.empty((P, n, 2))
        for i in range(P):
            b1[i], b2[i], _ = scene_H(n, n_in, 70 + i)
        seeds = np.arange(P, dtype=np.uint64) + 3
        monkeypatch.delenv("DGB200_FILTER32", raising=False)
        on = _cabi.homography_batch(b1, b2, 3.0, 0.999, 10000, 0, True, 0.0, seeds)
        monkeypatch.setenv("DGB200_FILTER32", "0")
        off = _cabi.homography_batch(b1, b2, 3.0, 0.999, 10000, 0, True, 0.0, seeds)
        monkeypatch.delenv("DGB200_FILTER32")
        for x, y in zip(on, off):
            assert np.array_equal(x, y)
        a = ref_oracle.find_homography_raw(b1[0], b2[0], 3.0, 0.999, 10000, seed=int(seeds[0]))
        _cmp(a, (on[0][0], on[1][0], on[2][0]), "H n=%d" % n)


def test_device_pointers_need_only_8_byte_alignment():
    """[n,2] inputs that are 8- but not 16-byte aligned (an offset view) take the scalar staging path."""
    import torch
    from pydegensac_b200 import _cabi
    dev = torch.device("cuda:0")
    P, N = 8, 1000
    b1, b2 = batch_F(P, N, 0.4, seed0=3)
    seeds = torch.arange(P, dtype=torch.int64, device=dev)
    res = []
    for shift in (0, 1):
        buf1 = torch.zeros(P * N * 2 + 2, dtype=torch.float64, device=dev)
        buf2 = torch.zeros(P * N * 2 + 2, dtype=torch.float64, device=dev)
        buf1[shift:shift + P * N * 2] = torch.from_numpy(b1.reshape(-1)).to(dev)
        buf2[shift:shift + P * N * 2] = torch.from_numpy(b2.reshape(-1)).to(dev)
        F = torch.zeros((P, 9), dtype=torch.float64, device=dev); m = torch.zeros((P, N), dtype=torch.uint8, device=dev)
        st = torch.zeros((P, 4), dtype=torch.int32, device=dev)
        _cabi.fundamental_batch_dev(buf1.data_ptr() + 8 * shift, buf2.data_ptr() + 8 * shift, P, N, 2, 1.0, 0.999, 2000, 0,
                                    True, 0.0, True, seeds.data_ptr(), F.data_ptr(), m.data_ptr(), st.data_ptr(), 0)
        torch.cuda.synchronize()
        res.append((F.cpu().numpy(), m.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_shared_reciprocal_division_is_ieee_exact():
    """h_pinvJ divides its eight entries by one denominator through a shared correctly rounded reciprocal + two FMA
    corrections; the quotients must be bit-identical to IEEE division (10^8 random operand pairs on the device)."""
    import ctypes
    from pydegensac_b200 import _cabi
    L = _cabi.lib()
    L.dgb200_debug_div_check.argtypes = [ctypes.c_ulonglong, ctypes.c_longlong, ctypes.POINTER(ctypes.c_ulonglong)]
    for seed in (1, 2):
        bad = ctypes.c_ulonglong(123)
        assert L.dgb200_debug_div_check(seed, 100_000_000, ctypes.byref(bad)) == 0
        assert bad.value == 0, "%d of 1e8 quotients differ from IEEE division" % bad.value


def test_large_and_extreme_sizes(ref_oracle):
    """N far beyond the benchmark sizes (rows longer than 8 residuals per thread, FP32 tile in the slab, queue sizes) and
    the smallest legal inputs."""
    from pydegensac_b200 import _cabi
    p1, p2, _ = scene_F(20000, 0.35, 11)
    a = ref_oracle.find_fundamental(p1, p2, 1.0, 0.999, 2000, seed=5)
    F, m, s = _cabi.fundamental_batch(p1, p2, 1.0, 0.999, 2000, 0, True, 0.0, True, [5])
    _cmp(a, (F[0], m[0], s[0]), "F n=20000")
    q1, q2, _ = scene_H(30000, 9000, 12)
    a = ref_oracle.find_homography_raw(q1, q2, 2.0, 0.999, 1000, seed=6)
    H, m, s = _cabi.homography_batch(q1, q2, 2.0, 0.999, 1000, 0, True, 0.0, [6])
    _cmp(a, (H[0], m[0], s[0]), "H n=30000")
    # 70000 correspondences: beyond the 16-bit packed counters of the two-threshold compaction
    r1, r2, _ = scene_F(70000, 0.5, 13)
    F, m, s = _cabi.fundamental_batch(r1, r2, 1.0, 0.99, 300, 0, True, 0.0, True, [7])
    a = ref_oracle.find_fundamental(r1, r2, 1.0, 0.99, 300, seed=7)
    _cmp(a, (F[0], m[0], s[0]), "F n=70000")
    # minimum sizes: the calls must succeed and agree between batch positions
    t1, t2, _ = scene_F(8, 1.0, 3)
    F8, m8, _ = _cabi.fundamental_batch(np.stack([t1, t1]), np.stack([t2, t2]), 1.0, 0.99, 200, 0, True, 0.0, True, [1, 1])
    assert np.array_equal(F8[0], F8[1]) and np.array_equal(m8[0], m8[1])
    h1, h2, _ = scene_H(4, 4, 3)
    H4, m4, _ = _cabi.homography_batch(np.stack([h1, h1]), np.stack([h2, h2]), 1.0, 0.99, 200, 0, True, 0.0, [1, 1])
    assert np.array_equal(H4[0], H4[1]) and np.array_equal(m4[0], m4[1])
