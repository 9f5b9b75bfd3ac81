"""GPU-resident global keypoint map and the frame-to-map flow against the oracle's
restatement of src/mymap.cpp / src/keypoint.cpp / updateMap (oracle.Map; order inside a block = insertion order)."""
import numpy as np
import pytest


def random_keypoints(n, seed, spread=40000.0, dup=0.15):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    xyz[:, 2] *= 0.2
    m = int(n * dup)                                   # clusters: neighbours within 800 mm and exact re-observations
    src = rng.integers(0, n, m)
    xyz[rng.integers(0, n, m)] = xyz[src] + rng.uniform(-600, 600, (m, 3)).astype(np.float32)
    xyz[rng.integers(0, n, m // 2)] = xyz[rng.integers(0, n, m // 2)]
    ratio = rng.choice(np.linspace(0.1, 1.0, 40).astype(np.float32), n)   # many equal saliencies (the rule uses <=)
    return xyz, ratio


def test_oracle_map_rules(oracle, synth):
    """hand cases of src/mymap.cpp:4-26 and src/keypoint.cpp:23-32 on the oracle itself"""
    d = synth.random_descriptors(6, seed=1)
    m = oracle.Map()
    m.add([[1234.9, -15.9, 7.0]], [0.5], d[:1])                      # snapped by truncation: (1230, -10, 0)
    xyz, desc = m.get([0, 0, 0], 100000.0)
    assert np.array_equal(xyz, [[1230.0, -10.0, 0.0]]) and np.array_equal(desc, d[:1])
    m.add([[1500.0, 0.0, 0.0]], [0.5], d[1:2])                      # within 800 mm, equal saliency: rejected (<=)
    assert len(m) == 1
    m.add([[1500.0, 0.0, 0.0]], [0.6], d[2:3])                      # more salient: inserted (the old one stays)
    assert len(m) == 2
    m.add([[1239.0, -19.0, 9.0]], [0.9], d[3:4])                    # same lattice position as the first: overwritten in place
    xyz, desc = m.get([0, 0, 0], 100000.0)
    assert len(m) == 2 and np.array_equal(desc, np.stack([d[3], d[2]]))
    m.add([[5100.0, 0.0, 0.0]], [0.1], d[4:5])                      # other 10 m block: no test across blocks
    assert len(m) == 3
    assert len(m.get([200000.0, 0, 0], 100000.0)[0]) == 0           # out of range


@pytest.mark.gpu
def test_gpu_map_matches_oracle(bshot, oracle, synth):
    with bshot.Context(0, 1024, 4096, 1 << 16) as ctx:
        ctx.gmap_create(1 << 16, 4096)
        om = oracle.Map()
        rng = np.random.default_rng(3)
        for frame in range(6):
            n = int(rng.integers(500, 3000))
            xyz, ratio = random_keypoints(n, 10 + frame)
            desc = synth.random_descriptors(n, seed=100 + frame)
            yaw = 0.05 * frame
            pose = None if frame == 0 else np.array([[np.cos(yaw), -np.sin(yaw), 0, 700.0 * frame], [np.sin(yaw), np.cos(yaw), 0, -55.5],
                                                     [0, 0, 1, 12.25]], np.float32)
            om.add(xyz, ratio, desc, pose)
            ctx.gmap_add(xyz, ratio, desc, pose)
            assert ctx.gmap_size() == (len(om), 0)
            for pos, r in (([0, 0, 0], 100000.0), ([12000.0, -3000.0, 500.0], 15000.0), ([-30000.0, 30000.0, 0.0], 4999.0)):
                xo, do = om.get(pos, r)
                xg, dg = ctx.gmap_get_keypoints(pos, r)
                assert np.array_equal(xg, xo) and np.array_equal(dg, do), (frame, pos, r, len(xo), len(xg))
        ctx.gmap_reset()
        assert ctx.gmap_size() == (0, 0)


@pytest.mark.gpu
def test_frame_to_map_flow(bshot, oracle, synth):
    """extract -> match against (map in range ++ previous frame) -> update map -> commit, three frames; the oracle replays
    the map and the matcher on the GPU's own keypoints / descriptors: target set and correspondences must be identical"""
    p = bshot.default_params(top_k=400)
    poses = [np.array([[1, 0, 0, 500.0 * k], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32) for k in range(3)]
    with bshot.Context(0, 65536, 400, 1 << 14) as ctx:
        ctx.gmap_create(1 << 14, 1024)
        om, prev = oracle.Map(), None
        for k in range(3):
            scan = synth.make_scan("hdl32e", k)[::2].copy()
            f = ctx.extract_frame(scan, p)
            assert len(f["bits"]) == 400
            if prev is not None:
                ref_pos = poses[k - 1][:, 3]
                r = ctx.match_frame_to_map(ref_pos, 100000.0, poses[k - 1])
                mx, md = om.get(ref_pos, 100000.0)
                R, T = poses[k - 1][:, :3], poses[k - 1][:, 3]
                tx = np.concatenate([mx, (prev["kp_xyz"] @ R.T + T).astype(np.float32)])
                td = np.concatenate([md, prev["bits"]])
                assert r["n_targets"] == len(td)
                assert np.allclose(r["target_xyz"], tx, atol=1e-2)
                m = oracle.match(f["bits"], td)
                assert np.array_equal(r["pairs"], oracle.mutual(m["left_idx"], m["right_idx"]))
                assert len(r["pairs"]) > 20
            ctx.gmap_update_from_frame(poses[k])
            om.add(f["kp_xyz"], f["seg_ratio"], f["bits"], poses[k])
            ctx.frame_commit()
            ctx.sync()
            assert ctx.gmap_size() == (len(om), 0)
            prev = f
