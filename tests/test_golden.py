"""Committed golden vectors (tests/golden/frontend_small.npz, made by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_small.npz"))
R, NN, K = 800.0, 60, 64


def test_oracle_reproduces_golden(oracle):
    c = oracle.Cloud(G["pts"])
    ratio = c.seg_ratio(R, NN, oracle.SR_CV, threads=1)
    assert np.array_equal(ratio, G["ratio"], equal_nan=True)
    idx, rat = oracle.select_keypoints(ratio, K, oracle.TIE_DETERMINISTIC)
    assert np.array_equal(idx, G["kp_idx"])
    d = c.compute_descriptors(G["pts"][idx], R, NN, oracle.MODE_REFERENCE, threads=1)
    assert np.array_equal(d["bits"], G["bits"])
    assert np.allclose(d["rf"], G["rf"], atol=1e-6, equal_nan=True)
    assert np.allclose(d["shot"], G["shot"], atol=1e-6, equal_nan=True)
    m = oracle.match(G["bits"], G["bits"][::-1].copy())
    assert np.array_equal(m["left_idx"], G["left_idx"]) and np.array_equal(m["right_idx"], G["right_idx"])


@pytest.mark.gpu
def test_gpu_matches_golden(bshot, synth):
    with bshot.Context(0, 4096, 256, 1024) as ctx:
        ctx.set_cloud(G["pts"])
        ratio = ctx.seg_ratio(R, NN, 0)
        ok = ~np.isnan(G["ratio"])
        assert np.array_equal(np.isnan(ratio), ~ok)
        assert np.array_equal(ratio[ok], G["ratio"][ok])          # bit-exact scores
        idx, _, _ = ctx.detect_keypoints(R, NN, 0, K)
        assert np.array_equal(idx, G["kp_idx"])                    # keypoint index equality
        ctx.set_keypoints(G["pts"][G["kp_idx"]])
        bits = ctx.compute_descriptors(bshot.default_params(top_k=K, kp_radius=R, kp_max_nn=NN, normal_radius=R,
                                                            normal_max_nn=NN, shot_radius=R))
        assert (synth.unpack_bits(bits) == synth.unpack_bits(G["bits"])).mean() >= 0.999
        rf, _ = ctx.compute_lrf(R)
        okr = ~np.isnan(G["rf"]).any(1)
        assert np.abs(rf[okr] - G["rf"][okr]).max() <= 1e-4
        m = ctx.match(G["bits"], G["bits"][::-1].copy())
        assert np.array_equal(m["left_idx"], G["left_idx"]) and np.array_equal(m["left_dist"], G["left_dist"])
        assert np.array_equal(m["right_idx"], G["right_idx"])
