"""Prints the measured parity figures of the SHOT / B-SHOT stage against the oracle (same keypoints, same normals):
max |LRF diff|, max |SHOT352 diff|, fraction of identical B-SHOT bits, for REFERENCE and FULL normals."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_bshot, load_synth, load_oracle
bs, synth, oracle = load_bshot(), load_synth(), load_oracle()
R = 3000.0
ctx = bs.Context(0, 131072, 16384, 16384)
for sensor, frame, K in (("hdl32e", 0, 600), ("hdl32e", 3, 2048), ("hdl64e", 0, 1024)):
    scan = synth.make_scan(sensor, frame)
    oc = oracle.Cloud(scan)
    ratio = oc.seg_ratio(R, 300, oracle.SR_CV)
    idx, _ = oracle.select_keypoints(ratio, K, oracle.TIE_DETERMINISTIC)
    kp = scan[idx]
    for mode in ("reference", "full"):
        if mode == "reference":
            od = oc.compute_descriptors(kp, R, 300, oracle.MODE_REFERENCE, want_normals=True)
            normals = od["normals"]
        else:
            normals = oc.normals(scan, 600.0, 60)
            shot, rf, nn, total = oc.shot(kp, normals, R)
            od = dict(shot=shot, rf=rf, bits=oracle.bshot(shot), sum_neighbours=total)
        ctx.reset(); ctx.set_cloud(scan); ctx.set_keypoints(kp); ctx.set_normals(normals)
        g = ctx.compute_shot(R, want_shot=True)
        ok = ~np.isnan(od["shot"]).any(1)
        bg, bo = synth.unpack_bits(g["bits"]), synth.unpack_bits(od["bits"])
        print(f"{sensor} f{frame} K={K} {mode:9s}: lrf max {np.abs(g['rf'][ok]-od['rf'][ok]).max():.2e}  shot max {np.abs(g['shot'][ok]-od['shot'][ok]).max():.2e}  "
              f"bits identical {(bg==bo).mean():.6f}  descriptors identical {(bg==bo).all(1).mean():.4f}  nn equal {g['sum_neighbours']==od['sum_neighbours']}")
