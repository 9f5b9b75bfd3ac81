"""The reference's per-frame loop (test/odometry_test.cpp:143-184) from raw laser returns to a pose, every stage through
the C ABI: Preprocessor::run -> extractKeypoints -> computeDescriptors -> featureMatching (map in range ++ previous frame,
mutual Hamming, RANSAC) -> evaluateEstimation (gate + ICP) -> poseEstimation -> updateMap.
The oracle replays the deterministic stages (map, matcher, RANSAC, gate, ICP) on the device's own keypoints and
descriptors: poses must be identical bit for bit; extraction parity itself is covered by test_frontend_gpu /
test_detector_edge_gpu, the preprocessor by test_preprocess.  The trajectory is also held against the synthetic truth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lasers_to_pose_four_frames(bshot, oracle, synth):
    p = bshot.default_params(top_k=600)
    ident = np.eye(4, dtype=np.float32)
    with bshot.Context(0, 131072, 1024, 1 << 15) as ctx:
        ctx.gmap_create(1 << 15, 4096)
        om = oracle.Map()
        pose_ref, prev, poses = ident, None, []
        for k in range(4):
            L = synth.make_lasers("hdl32e", k)
            cloud = ctx.preprocess(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"])
            if oracle.ref_lib() is not None:
                assert np.array_equal(cloud, oracle.ref_preprocess(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"]))
            f = ctx.extract_frame(cloud, p)
            assert len(f["bits"]) == 600
            if prev is None:   # initial frame: the target is the frame itself (src/lidar_odometry.cpp:187-194)
                pairs = ctx.match_mutual(f["bits"], f["bits"])[0]
                tgt_xyz, o_pairs = f["kp_xyz"], None
                m = oracle.match(f["bits"], f["bits"])
                o_pairs, o_tgt = oracle.mutual(m["left_idx"], m["right_idx"]), f["kp_xyz"]
            else:
                r = ctx.match_frame_to_map(pose_ref[:3, 3], 100000.0, pose_ref[:3])
                pairs, tgt_xyz = r["pairs"], r["target_xyz"]
                mx, md = om.get(pose_ref[:3, 3], 100000.0)
                o_tgt = np.concatenate([mx, (prev["kp_xyz"] @ pose_ref[:3, :3].T + pose_ref[:3, 3]).astype(np.float32)])
                m = oracle.match(f["bits"], np.concatenate([md, prev["bits"]]))
                o_pairs = oracle.mutual(m["left_idx"], m["right_idx"])
            assert np.array_equal(pairs, o_pairs)
            assert np.allclose(tgt_xyz, o_tgt, atol=1e-2)
            rs = ctx.ransac(f["kp_xyz"], tgt_xyz, pairs)
            ev = ctx.evaluate_estimation(rs["transform"], pose_ref, len(rs["pairs"]), f["kp_xyz"], tgt_xyz, run_icp=True)
            ors = oracle.ransac(f["kp_xyz"], tgt_xyz, pairs)
            oev = oracle.evaluate_estimation(ors["transform"], pose_ref, len(ors["pairs"]), f["kp_xyz"], tgt_xyz, run_icp=True)
            assert np.array_equal(rs["pairs"], ors["pairs"]) and np.array_equal(rs["transform"], ors["transform"])
            assert ev["should_update_map"] == oev["should_update_map"]
            assert np.array_equal(ev["T_best"].view(np.uint32), oev["T_best"].view(np.uint32))
            pose = ev["T_best"]                                   # poseEstimation: src_->setPose(T_best_)
            ctx.gmap_update_from_frame(pose[:3])                  # updateMap: every keypoint, moved by T_best_
            om.add(f["kp_xyz"], f["seg_ratio"], f["bits"], pose[:3])
            ctx.frame_commit()
            ctx.sync()
            assert ctx.gmap_size()[0] == len(om)
            poses.append(pose.copy())
            pose_ref, prev = pose, f
    # the sensor moves 500 mm a frame along x without turning
    for k, T in enumerate(poses):
        assert np.allclose(T[:3, :3], np.eye(3), atol=0.02), (k, T)
        assert abs(T[0, 3] - 500.0 * k) < 350.0 and abs(T[1, 3]) < 350.0 and abs(T[2, 3]) < 350.0, (k, T[:3, 3])


def test_extract_scan_equals_preprocess_then_extract_frame(bshot, synth):
    """bshot_extract_scan keeps the preprocessed cloud on the device: same cloud, keypoints and descriptors as the two calls"""
    p = bshot.default_params(top_k=600)
    L = synth.make_lasers("hdl32e", 2)
    with bshot.Context(0, 131072, 1024, 1024) as a, bshot.Context(0, 131072, 1024, 1024) as b:
        cloud = a.preprocess(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"])
        fa = a.extract_frame(cloud, p)
        fb = b.extract_scan(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"], p)
        assert np.array_equal(fb["cloud"], cloud) and fb["n_points"] == len(cloud)
        for key in ("kp_idx", "kp_xyz", "seg_ratio", "bits"):
            assert np.array_equal(fa[key], fb[key]), key
        fc = b.extract_scan(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"], p, want_cloud=False)
        assert fc["cloud"] is None and np.array_equal(fc["bits"], fa["bits"])
        with pytest.raises(bshot.BshotError):   # more points than the context was created for
            with bshot.Context(0, 1024, 64, 64) as small:
                small.extract_scan(L["azimuth"], L["vertical"], L["distance"], L["ring_deg"], bshot.default_params(top_k=64))
