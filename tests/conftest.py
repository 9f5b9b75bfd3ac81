"""Shared test plumbing: loads the product binding (bshot_b200), the oracle binding (pyoracle)
and the synthetic-scan generator.  GPU tests are marked `gpu` and go through the C ABI."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "b-shot-slam_b200")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_bshot():
    return _load("bshot_b200", os.path.join(PKG, "__init__.py"))


def load_synth():
    return _load("bshot_b200_synth", os.path.join(PKG, "synth.py"))


def load_sharded():
    return _load("bshot_b200_sharded", os.path.join(PKG, "sharded.py"))


def load_oracle():
    return _load("pyoracle", os.path.join(ROOT, "oracle", "pyoracle.py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def bshot():
    return load_bshot()


@pytest.fixture(scope="session")
def synth():
    return load_synth()


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


@pytest.fixture(scope="session")
def gpu_ctx(bshot):
    ctx = bshot.Context(device=0, max_points=131072, max_keypoints=16384, max_targets=1 << 21)
    yield ctx
    ctx.close()
