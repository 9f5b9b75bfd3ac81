"""Generates tests/golden/preprocess_pin.npz from oracle/_ref/libbshot_ref.so, i.e. from the REFERENCE's own
src/preprocess.cpp compiled unchanged (oracle/Makefile `ref`; Eigen's Vector3f / velodyne::Laser / CV_PI ->
oracle/pre_stub/pre_stub.h).  Run in the build container (needs /root/reference):
    python tests/golden/make_preprocess_pin.py
The inputs are regenerated from the seeded synthetic rotation (b-shot-slam_b200/synth.py make_lasers); their checksum is
stored so that a drift of the generator is noticed.  The vectors pin the GPU preprocessor on any machine."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_oracle, load_synth  # noqa: E402

CASES = {"hdl32e": dict(sensor="hdl32e", frame=0, firings=300, start_deg=340.0), "hdl64e": dict(sensor="hdl64e", frame=2, firings=120, start_deg=80.0)}


def lasers_crc(L):
    return zlib.crc32(L["azimuth"].tobytes() + L["vertical"].tobytes() + L["distance"].tobytes())


def main():
    oracle, synth = load_oracle(), load_synth()
    assert oracle.ref_lib() is not None, "needs /root/reference"
    out = {}
    for name, kw in CASES.items():
        L = synth.make_lasers(**kw)
        out[name + "_crc"] = np.uint32(lasers_crc(L))
        out[name + "_xyz"] = oracle.ref_preprocess(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"])
        print(name, L["distance"].size, "returns ->", out[name + "_xyz"].shape[0], "points")
    np.savez_compressed(os.path.join(HERE, "preprocess_pin.npz"), **out)


if __name__ == "__main__":
    main()
