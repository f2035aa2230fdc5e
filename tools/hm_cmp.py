import sys
import numpy as np
import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
sys.path.insert(0, ".")
from tests import util
from tests.test_tracker_edge_gpu import _feed, JOINT
for (w, h) in ((640, 480), (320, 240)):
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    p0 = pose0.astype(np.float32)
    mk = lambda mode: ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=mode)
    plain, fused, dev = mk(RO.EF_SOLVE_HOST), mk(RO.EF_SOLVE_HOST), mk(RO.EF_SOLVE_DEVICE)
    plain.set_option(RO.EF_OPT_HOST_FUSED, 0)
    for name, m in (("joint", JOINT), ("rgb_only", dict(JOINT, rgbOnly=True)), ("rgb_only_l0", dict(JOINT, rgbOnly=True, pyramid=False, fastOdom=True))):
        res = []
        for t in (plain, fused, dev):
            _feed(t, p0, f0, f1)
            tt, R = t.getIncrementalTransformation(p0[:3, 3], p0[:3, :3], **m)
            res.append((tt, t.se3_iterations, t.lastRGBCount, t.lastRGBError))
        d = lambda a, b: float(np.abs(res[a][0] - res[b][0]).max())
        print(f"{w}x{h} {name:12s} |plain-fused| {d(0,1):.2e}  |plain-device| {d(0,2):.2e}  |fused-device| {d(1,2):.2e}  iters {res[0][1]} {res[1][1]} {res[2][1]}  rgbcount {res[0][2]:.0f} {res[1][2]:.0f} {res[2][2]:.0f}  true {np.abs(res[0][0]-pose1[:3,3]).max():.2e}")
