"""Generates tests/golden/golden_laf_v1.npz from the UNMODIFIED reference (oracle/_ref, Philox replay stream):
LAF-consistency gate cases ([N,6] inputs, laf_coef > 0) for both drivers.  Run in the build container:

    python tests/golden/make_golden_laf.py

Cases are picked so that the gate changes the outcome for most of them (the script prints whether it did)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from pydegensac_b200.scenes import scene_F_laf, scene_H_laf  # noqa: E402

CASES = [
    ("F", dict(n=600, inlier_ratio=0.6, seed=301, jitter=0.6, plane_frac=0.0), dict(px_th=1.0, conf=0.999, max_iters=3000, error_type=0, sym_check=True, laf_coef=1.0, degen_check=True, seed=21)),
    ("F", dict(n=1000, inlier_ratio=0.4, seed=302, jitter=1.5, plane_frac=0.0), dict(px_th=1.0, conf=0.999, max_iters=3000, error_type=1, sym_check=False, laf_coef=2.0, degen_check=True, seed=22)),
    ("F", dict(n=300, inlier_ratio=0.8, seed=303, jitter=0.2, plane_frac=0.6), dict(px_th=1.0, conf=0.999, max_iters=1000, error_type=0, sym_check=True, laf_coef=0.5, degen_check=True, seed=23)),
    ("F", dict(n=2000, inlier_ratio=0.3, seed=304, jitter=0.6, plane_frac=0.0), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, laf_coef=5.0, degen_check=True, seed=24)),
    # three F scenes on which the gate changes the final mask (found by a seed search against the reference)
    ("F", dict(n=1000, inlier_ratio=0.4, seed=42210, jitter=3.0, plane_frac=0.0), dict(px_th=1.0, conf=0.999, max_iters=1500, error_type=0, sym_check=True, laf_coef=2.0, degen_check=True, seed=42210)),
    ("F", dict(n=300, inlier_ratio=0.6, seed=53487, jitter=1.5, plane_frac=0.0), dict(px_th=1.0, conf=0.999, max_iters=1500, error_type=0, sym_check=True, laf_coef=2.0, degen_check=True, seed=53487)),
    ("F", dict(n=300, inlier_ratio=0.4, seed=498, jitter=1.5, plane_frac=0.0), dict(px_th=1.0, conf=0.999, max_iters=500, error_type=1, sym_check=True, laf_coef=0.5, degen_check=True, seed=498)),
    ("H", dict(n=800, n_in=480, seed=311, jitter=0.6), dict(px_th=3.0, conf=0.999, max_iters=3000, error_type=0, sym_check=True, laf_coef=2.0, seed=31)),
    ("H", dict(n=800, n_in=320, seed=312, jitter=1.5), dict(px_th=1.0, conf=0.999, max_iters=3000, error_type=1, sym_check=True, laf_coef=5.0, seed=32)),
    ("H", dict(n=300, n_in=240, seed=313, jitter=0.2), dict(px_th=3.0, conf=0.999, max_iters=1000, error_type=2, sym_check=False, laf_coef=1.0, seed=33)),
    ("H", dict(n=300, n_in=180, seed=314, jitter=0.6), dict(px_th=1.0, conf=0.999, max_iters=1000, error_type=3, sym_check=True, laf_coef=20.0, seed=34)),
    ("H", dict(n=5000, n_in=1500, seed=315, jitter=0.6), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=4, sym_check=True, laf_coef=2.0, seed=35)),
]


def checksum(a, b):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes() + np.ascontiguousarray(b).tobytes()).hexdigest()[:16]


def main():
    out = {}
    meta = []
    for i, (kind, sargs, kw) in enumerate(CASES):
        kw0 = dict(kw); kw0["laf_coef"] = 0.0
        if kind == "F":
            p1, p2, _ = scene_F_laf(**sargs)
            M, mask, stats = ref.find_fundamental(p1, p2, **kw)
            m0 = ref.find_fundamental(p1, p2, **kw0)[1]
        else:
            p1, p2, _ = scene_H_laf(**sargs)
            M, mask, stats = ref.find_homography_raw(p1, p2, **kw)
            m0 = ref.find_homography_raw(p1, p2, **kw0)[1]
        out["model_%d" % i] = M
        out["mask_%d" % i] = mask
        out["stats_%d" % i] = stats
        meta.append(dict(kind=kind, scene=sargs, call=kw, input_sha=checksum(p1, p2), gate_changed_mask=bool(not np.array_equal(mask, m0))))
        print(i, kind, stats, int(mask.sum()), "gate changed the mask:", not np.array_equal(mask, m0))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_laf_v1.npz"), **out)


if __name__ == "__main__":
    main()
