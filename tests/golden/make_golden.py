"""Regenerates tests/golden/frontend_small.npz.

The reference (C++/PCL) cannot run here and ships no fixtures (parity unpinned), so these
vectors are ORACLE outputs on a tiny seeded cloud, committed to catch silent changes of the oracle and
to give the GPU tests a size-independent fixed case.  They are cross-checked against the independent
numpy restatement (tests/shot_numpy.py) before being written.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_oracle, load_synth  # noqa: E402
from shot_numpy import lrf_numpy, shot_numpy  # noqa: E402


def cloud():
    rng = np.random.default_rng(20260118)
    a = rng.uniform(-2000, 2000, (900, 3))
    a[:, 2] = 0.2 * a[:, 0] - 0.1 * a[:, 1] + rng.normal(0, 25, 900)
    b = rng.uniform(-2000, 2000, (600, 3))
    b[:, 1] = 700 + rng.normal(0, 20, 600)
    return np.concatenate([a, b]).astype(np.float32)


def main():
    o, synth = load_oracle(), load_synth()
    pts = cloud()
    c = o.Cloud(pts)
    ratio = c.seg_ratio(800.0, 60, o.SR_CV, threads=1)
    idx, rat = o.select_keypoints(ratio, 64, o.TIE_DETERMINISTIC)
    kp = pts[idx]
    d = c.compute_descriptors(kp, 800.0, 60, o.MODE_REFERENCE, threads=1, want_normals=True)
    for i in range(0, 64, 8):      # cross-check before committing
        rf, _ = lrf_numpy(pts, kp[i], 800.0)
        assert np.allclose(rf, d["rf"][i], atol=1e-5, equal_nan=True)
        assert np.allclose(shot_numpy(pts, kp[i], 800.0, d["normals"], d["rf"][i]), d["shot"][i], atol=2e-6, equal_nan=True)
    m = o.match(d["bits"], d["bits"][::-1].copy())
    np.savez_compressed(os.path.join(HERE, "frontend_small.npz"), pts=pts, ratio=ratio, kp_idx=idx, kp_ratio=rat,
                        rf=d["rf"], shot=d["shot"].astype(np.float32), bits=d["bits"], normals_k=d["normals"][:64],
                        left_idx=m["left_idx"], left_dist=m["left_dist"], right_idx=m["right_idx"])
    print("written", os.path.join(HERE, "frontend_small.npz"))


if __name__ == "__main__":
    main()
