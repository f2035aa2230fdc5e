"""Diagnostic: per-frame deviation of the product from the reference's CUDA tracker next to the reference's OWN
spread when only its launch shape changes (its float sums depend on (threads, blocks), reduce.cu:90-255).
    python tools/parity_spread.py [--frames 31]"""
import argparse
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from oracle import oracle as O
from tests import util
from tests.test_trajectory_parity_gpu import _render

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=31)
args = ap.parse_args()

ALT = [(256, 96), (96, 148), (512, 32)]


def ref_with(w, h, K, cfg):
    r = O.OracleTracker(w, h, K.cx, K.cy, K.fx, K.fy, impl="ref")
    if cfg:
        t, b = cfg
        r.lib.efr_tracker_set_config(r.t, t, b, t, b, t, b, t, b)
    return r


for (w, h), so3 in (((640, 480), False), ((1280, 720), True)):
    K, poses, dev, host = _render(w, h, args.frames)
    posef = poses.astype(np.float32)
    kw = dict(rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=so3)
    refs = [ref_with(w, h, K, None)] + [ref_with(w, h, K, c) for c in ALT]
    prods = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=m) for m in (RO.EF_SOLVE_DEVICE, RO.EF_SOLVE_HOST)]
    if so3:
        for r in refs:
            r.init_first_rgb(host[0]["rgba"])
        for p in prods:
            p.initFirstRGB(host[0]["rgba"])
    print(f"== {w}x{h} so3={so3}: |dt| (um) / rot (urad) vs reference at GPUConfig defaults; ref-alt = reference at {ALT}")
    worst = np.zeros((2 + len(ALT), 2))
    for k in range(1, args.frames):
        p = posef[k - 1]
        out = []
        for r in refs:
            r.init_icp_model(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
            r.init_rgb_model(host[k - 1]["rgba"])
            r.init_icp_depth(host[k]["depth"], 20.0)
            r.init_rgb(host[k]["rgba"])
            out.append(r.get_incremental_transformation(p[:3, 3], p[:3, :3], **kw))
        res = []
        for pr in prods:
            pr.initICPModel(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
            pr.initRGBModel(host[k - 1]["rgba"])
            pr.initICP(host[k]["depth"], 20.0)
            pr.initRGB(host[k]["rgba"])
            res.append(pr.getIncrementalTransformation(p[:3, 3], p[:3, :3], False, 10.0, True, False, so3))
        t0, R0, st0 = out[0]
        row = []
        for (t, R) in res:
            row.append((float(np.abs(t - t0).max()) * 1e6, util.rot_err(R, R0) * 1e6))
        for (t, R, st) in out[1:]:
            row.append((float(np.abs(t - t0).max()) * 1e6, util.rot_err(R, R0) * 1e6))
        worst = np.maximum(worst, np.array(row))
        A0 = st0["last_A"]; b0 = st0["last_b"]
        dA = [np.linalg.norm(pr.lastA - A0) / np.linalg.norm(A0) for pr in prods] + [np.linalg.norm(o[2]["last_A"] - A0) / np.linalg.norm(A0) for o in out[1:]]
        db = [np.linalg.norm(pr.lastb - b0) for pr in prods] + [np.linalg.norm(o[2]["last_b"] - b0) for o in out[1:]]
        dc = [pr.lastICPCount - st0["last_icp_count"] for pr in prods] + [o[2]["last_icp_count"] - st0["last_icp_count"] for o in out[1:]]
        dr = [pr.lastRGBCount - st0["last_rgb_count"] for pr in prods] + [o[2]["last_rgb_count"] - st0["last_rgb_count"] for o in out[1:]]
        print(f"k={k:2d} " + " ".join(f"{a:5.1f}/{b:5.1f}" for a, b in row) + "  dA " + " ".join(f"{x:.1e}" for x in dA) + f"  |b|={np.linalg.norm(b0):.2e} db " +
              " ".join(f"{x:.1e}" for x in db) + "  dICP " + " ".join(f"{int(x):d}" for x in dc) + "  dRGB " + " ".join(f"{int(x):d}" for x in dr) + f"  |A|2={np.linalg.norm(A0, 2):.2e}")
    print("worst   " + " ".join(f"{a:5.1f}/{b:5.1f}" for a, b in worst), " columns: device-solve, host-solve, ref-alt x", len(ALT))
    for x in refs + prods:
        x.close()
