"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref) driven by the Philox replay
stream.  Run in the build container (needs /root/reference to have been compiled by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Each case stores the scene parameters (inputs are regenerated from pydegensac_b200.scenes, also stored as a
checksum), the call arguments and the reference outputs (model, mask, stats)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from pydegensac_b200.scenes import scene_F, scene_H  # noqa: E402

CASES = [
    # kind, scene args, call kwargs
    ("F", dict(n=2000, inlier_ratio=0.30, seed=0, plane_frac=0.0), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, degen_check=True, seed=0)),
    ("F", dict(n=2000, inlier_ratio=0.30, seed=0, plane_frac=0.0), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, degen_check=True, seed=1)),
    ("F", dict(n=2000, inlier_ratio=0.30, seed=1, plane_frac=0.0), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, degen_check=False, seed=2)),
    ("F", dict(n=2000, inlier_ratio=0.30, seed=0, plane_frac=0.8), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, degen_check=True, seed=0)),
    ("F", dict(n=2000, inlier_ratio=0.30, seed=0, plane_frac=0.8), dict(px_th=1.0, conf=0.9999, max_iters=10000, error_type=0, sym_check=True, degen_check=True, seed=3)),
    ("F", dict(n=500, inlier_ratio=0.5, seed=7, plane_frac=0.0), dict(px_th=0.5, conf=0.999, max_iters=2000, error_type=1, sym_check=True, degen_check=True, seed=5)),
    ("F", dict(n=500, inlier_ratio=0.5, seed=7, plane_frac=0.5), dict(px_th=0.5, conf=0.999, max_iters=2000, error_type=0, sym_check=False, degen_check=True, seed=6)),
    ("F", dict(n=300, inlier_ratio=0.8, seed=9, plane_frac=0.0), dict(px_th=1.0, conf=0.99, max_iters=1000, error_type=0, sym_check=True, degen_check=True, seed=7)),
    ("F", dict(n=100, inlier_ratio=0.5, seed=11, plane_frac=0.9), dict(px_th=1.0, conf=0.9, max_iters=100, error_type=1, sym_check=True, degen_check=True, seed=8)),
    ("F", dict(n=40, inlier_ratio=1.0, seed=12, plane_frac=0.0), dict(px_th=3.0, conf=0.9999, max_iters=49, error_type=0, sym_check=True, degen_check=True, seed=9)),
    ("F", dict(n=12, inlier_ratio=1.0, seed=13, plane_frac=0.0), dict(px_th=1.0, conf=0.99, max_iters=200, error_type=0, sym_check=True, degen_check=False, seed=10)),
    ("F", dict(n=8, inlier_ratio=1.0, seed=14, plane_frac=0.0), dict(px_th=1.0, conf=0.99, max_iters=60, error_type=0, sym_check=True, degen_check=True, seed=11)),
    ("H", dict(n=5000, n_in=1500, seed=0), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=0, sym_check=True, seed=0)),
    ("H", dict(n=5000, n_in=1500, seed=0), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=1, sym_check=True, seed=1)),
    ("H", dict(n=5000, n_in=1500, seed=1), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=2, sym_check=True, seed=2)),
    ("H", dict(n=5000, n_in=1500, seed=2), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=3, sym_check=True, seed=3)),
    ("H", dict(n=5000, n_in=1500, seed=0), dict(px_th=3.0, conf=0.999, max_iters=10000, error_type=4, sym_check=True, seed=0)),
    ("H", dict(n=811, n_in=110, seed=5), dict(px_th=4.0, conf=0.99, max_iters=2000, error_type=0, sym_check=True, seed=4)),
    ("H", dict(n=300, n_in=200, seed=6), dict(px_th=1.0, conf=0.999, max_iters=500, error_type=0, sym_check=False, seed=5)),
    ("H", dict(n=50, n_in=25, seed=7), dict(px_th=1.0, conf=0.99, max_iters=100, error_type=4, sym_check=True, seed=6)),
    ("H", dict(n=20, n_in=20, seed=8), dict(px_th=2.0, conf=0.99, max_iters=49, error_type=0, sym_check=True, seed=7)),
    ("H", dict(n=4, n_in=4, seed=9), dict(px_th=2.0, conf=0.99, max_iters=60, error_type=0, sym_check=True, seed=8)),
]


def checksum(a, b):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes() + np.ascontiguousarray(b).tobytes()).hexdigest()[:16]


def main():
    out = {}
    meta = []
    for i, (kind, sargs, kw) in enumerate(CASES):
        if kind == "F":
            p1, p2, _ = scene_F(**sargs)
            M, mask, stats = ref.find_fundamental(p1, p2, **kw)
        else:
            p1, p2, _ = scene_H(**sargs)
            M, mask, stats = ref.find_homography_raw(p1, p2, **kw)
        out["model_%d" % i] = M
        out["mask_%d" % i] = mask
        out["stats_%d" % i] = stats
        meta.append(dict(kind=kind, scene=sargs, call=kw, input_sha=checksum(p1, p2)))
        print(i, kind, sargs, kw, stats, int(mask.sum()))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)


if __name__ == "__main__":
    main()
