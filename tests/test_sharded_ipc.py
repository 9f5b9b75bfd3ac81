"""The multi-rank exchange behind the C ABI (bshot_comm_* / bshot_match_map_sharded, include/bshot_b200.h): a C++
host -- like the reference (include/lidar_odometry.h:52-72) -- runs sharded frame-to-map matching without Python.
 * tests/sharded_ipc_test.cpp: one PROCESS per rank (fork before CUDA), handles over pipes, CUDA IPC peer memory;
 * in-process variant: two contexts on one GPU connected with device pointers (bshot_comm_import_ptrs).
Both must reproduce the single-GPU records bit for bit (first-minimum tie-break across shards, rq of the winner)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "b-shot-slam_b200")


@pytest.fixture(scope="module")
def ipc_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ipc") / "sharded_ipc_test")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "sharded_ipc_test.cpp"), "-o", out, "-L", PKG, "-lbshot_b200",
                           f"-Wl,-rpath,{PKG}"])
    return out


def test_driver_compiles_and_fails_loudly_without_gpu(ipc_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([ipc_binary, "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,T,Q,matcher", [(2, 6000, 700, None), (3, 5000, 257, None), (4, 40000, 1024, None), (2, 6000, 700, "2")],
                         ids=["2-ranks", "3-ranks", "4-ranks", "2-ranks-tensor-core-pipelined"])
def test_multi_process_sharded_match(ipc_binary, nranks, T, Q, matcher):
    import torch
    env = dict(os.environ, BSHOT_TEST_NDEV=str(torch.cuda.device_count()))
    if matcher is not None:
        env["BSHOT_MATCH_TC"] = matcher      # every search of the call on the tensor-core pipeline, whatever its size
    r = subprocess.run([ipc_binary, str(nranks), str(T), str(Q)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "all ranks ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_in_process_ranks_with_device_pointers():
    """ranks that live in ONE process (several contexts, device pointers instead of IPC handles).  The consumer kernels
    spin on flags that the other ranks' kernels release, so the streams must not share a hardware queue: the case runs in
    a fresh interpreter with CUDA_DEVICE_MAX_CONNECTIONS=32 (tests/ipc_inprocess_case.py); bshot_comm_create loads the
    kernels of the call up front, so lazy module loading cannot stall a producer behind a spinning consumer"""
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    r = subprocess.run([os.sys.executable, os.path.join(ROOT, "tests", "ipc_inprocess_case.py")], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0 and "in-process ranks ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_missing_rank_is_reported(bshot, synth):
    """a rank that never calls: the others give up after the bounded wait and the next check fails loudly"""
    T, Q = 2000, 64
    tmap = synth.random_descriptors(T, seed=3)
    q = synth.random_descriptors(Q, seed=4)
    a, b = bshot.Context(0, 256, 256, T), bshot.Context(0, 256, 256, T)
    try:
        a.map_append(tmap[:1000]); b.map_append(tmap[1000:])
        a.comm_create(0, 2, Q); b.comm_create(1, 2, Q)
        regions = [a.comm_region()[0], b.comm_region()[0]]
        a.comm_import_ptrs(regions); b.comm_import_ptrs(regions)
        with pytest.raises(bshot.BshotError, match="did not arrive"):
            a.match_map_sharded(q, 0)                           # rank 1 never shows up
    finally:
        a.close(); b.close()
