"""Generates tests/golden/ref_pin.npz from oracle/_ref/libbshot_ref.so, i.e. from the REFERENCE's own
include/bshot_bits.h compiled unchanged (oracle/ref_shim.cpp; PCL calls -> oracle/pcl_stub).  Run in the build
container (needs /root/reference):  python tests/golden/make_ref_pin.py
The vectors pin, on any machine: orc_bshot / orc_match / orc_mutual / orc_compute_descriptors (CPU tests) and
the GPU binarise / match / descriptor kernels (tests -m gpu) to reference-compiled code."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_oracle, load_synth  # noqa: E402


def shot_vectors(seed=11, n=256):
    """SHOT-like rows + every edge the binarisation rule has (zeros, lone spikes at the 0.9 boundary, negative
    and NaN entries, exact ties)"""
    rng = np.random.default_rng(seed)
    s = rng.random((n, 352)).astype(np.float32) ** 6          # mostly small with a few dominant bins
    s[rng.random((n, 352)) < 0.55] = 0.0                     # SHOT histograms are sparse
    s /= np.maximum(np.sqrt((s.astype(np.float64) ** 2).sum(1, keepdims=True)), 1e-12).astype(np.float32)
    s[0] = 0.0
    s[1] = np.nan
    s[2, ::4] = 0.9; s[2, 1::4] = 0.1; s[2, 2::4] = 0.0; s[2, 3::4] = 0.0     # v0 == 0.9*sum boundary (not >)
    s[3, ::4] = 0.91; s[3, 1::4] = 0.09; s[3, 2::4] = 0; s[3, 3::4] = 0
    s[4] = 0.25                                                                    # all equal -> 1111
    s[5, ::4] = -1.0; s[5, 1::4] = 1.0; s[5, 2::4] = 0.5; s[5, 3::4] = 0.0     # negative entries
    s[6, :176] = np.float32(1e-38); s[6, 176:] = np.float32(3e38)                # denormal-ish / huge
    s[7, ::7] = np.nan                                                             # scattered NaN
    for r in range(8, 72):                                                         # every pair/triple pattern near 0.9
        g = rng.random((88, 4)).astype(np.float32)
        k = rng.integers(1, 4, 88)
        for j in range(88):
            idx = rng.permutation(4)[:k[j]]
            rest = np.setdiff1d(np.arange(4), idx)
            g[j, idx] = rng.uniform(0.2, 1.0, k[j])
            tot = g[j, idx].sum()
            g[j, rest] = (tot / 9.0) * rng.uniform(0.9, 1.1) / max(len(rest), 1)
        s[r] = g.reshape(-1)
    return s


def descriptors(seed, n, dup_from=None):
    rng = np.random.default_rng(seed)
    bits = rng.random((n, 352)) < rng.uniform(0.05, 0.5, (n, 1))
    if dup_from is not None:                                    # exact duplicates and near duplicates -> ties
        m = min(n // 4, len(dup_from))
        bits[:m] = dup_from[:m]
        flip = rng.integers(0, 352, m // 2)
        bits[np.arange(m // 2), flip] ^= True
    return bits


def main():
    oracle, synth = load_oracle(), load_synth()
    assert oracle.ref_lib() is not None, "oracle/_ref/libbshot_ref.so could not be built (needs /root/reference)"
    out = {}
    s = shot_vectors()
    out["shot"] = s
    out["shot_bits"] = oracle.ref_bshot(s)
    tb = descriptors(21, 320)
    qb = descriptors(22, 200, dup_from=tb)
    tb[300:] = tb[:20]                                          # duplicate targets: first minimum must win
    q, t = synth.pack_bits(qb), synth.pack_bits(tb)
    m = oracle.ref_feature_matching(q, t)
    out.update(match_q=q, match_t=t, match_left=m["left_idx"], match_right=m["right_idx"], match_pairs=m["pairs"])
    # reference flow over two frames with a persistent `cb` (second frame has FEWER keypoints: stale normals quirk)
    cb = oracle.RefCb()
    f0 = synth.make_scan("hdl32e", 0)[::6].copy()
    f1 = synth.make_scan("hdl32e", 1)[::6].copy()
    kp0 = f0[np.random.default_rng(31).permutation(len(f0))[:96]]
    kp1 = f1[np.random.default_rng(32).permutation(len(f1))[:64]]
    d0 = cb.compute_descriptors(f0, kp0, 3000.0)
    d1 = cb.compute_descriptors(f1, kp1, 3000.0)
    out.update(flow_stride=np.int32(6), flow_kp0=kp0, flow_kp1=kp1, flow_bits0=d0["bits"], flow_bits1=d1["bits"],
               flow_rf0=d0["rf"], flow_normals0_head=d0["normals"][:128], flow_normals1_head=d1["normals"][:128],
               flow_shot0=d0["shot"][:16])
    np.savez_compressed(os.path.join(HERE, "ref_pin.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
