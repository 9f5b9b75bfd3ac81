"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/bshot_b200.h declares, and fails LOUDLY (no CPU fallback) when no B200 is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "bshot_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(bshot_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_are_exported(bshot):
    lib = bshot.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bshot_b200.h but not exported"
    assert sorted(bshot.EXPORTS) == names


def test_version_and_defaults(bshot):
    assert bshot.lib().bshot_version() == 100
    p = bshot.default_params()
    # the reference's literals (src/lidar_odometry.cpp:68,70,138-142,174-175; include/bshot_bits.h:68)
    assert (p.kp_radius, p.kp_max_nn, p.sr_type, p.top_k) == (3000.0, 300, 0, 600)
    assert (p.normal_radius, p.normal_max_nn, p.normals_mode, p.shot_radius) == (3000.0, 300, 0, 3000.0)


def test_cand_record_layout(bshot):
    assert bshot.CAND_DTYPE.itemsize == 24
    c = np.zeros(2, bshot.CAND_DTYPE)
    c["k1"] = [(5 << 32) | 77, 0xFFFFFFFFFFFFFFFF]
    c["k2"] = [(9 << 32) | 3, 0xFFFFFFFFFFFFFFFF]
    c["rq"] = [4, 0xFFFFFFFF]
    u = bshot.unpack_cands(c)
    assert list(u["idx1"]) == [77, -1] and list(u["dist1"]) == [5, -1]
    assert list(u["idx2"]) == [3, -1] and list(u["rq"]) == [4, -1]


def test_no_gpu_fails_loudly(bshot):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bshot.BshotError) as e:
        bshot.Context(0, 1024, 64, 64)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    # null-context calls report errors instead of crashing
    assert bshot.lib().bshot_ctx_sync(None) != 0
    assert b"null context" in bshot.lib().bshot_last_error()


def test_product_never_touches_the_oracle():
    """the product package must not import / link anything under oracle/"""
    pkg = os.path.join(ROOT, "b-shot-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "bshot_oracle" not in src and "oracle/" not in src, (dirpath, f)
    so = os.path.join(pkg, "libbshot_b200.so")
    if os.path.exists(so):
        import subprocess
        out = subprocess.check_output(["ldd", so]).decode()
        assert "oracle" not in out
