"""Frame-by-frame parity over consecutive frames of BASELINE.json's trajectories (configs[1] and configs[2]) against the
reference's own CUDA kernels under the restated host loop (oracle/_ref), with the tracker state CARRIED across frames
(the SO(3) pre-alignment swaps its image pyramids at the end of every call, RGBDOdometry.cpp:593-599 -- nothing is re-fed).

Open-loop protocol of SURVEY.md 8(d): model maps of frame k are ray-cast at the ground-truth pose k-1 and the tracker is
called with that pose as the prior.  Tolerances (BASELINE.json north_star): pose 1e-5 m / 1e-5 rad per frame; JtJ | Jtr
1e-4 norm-relative -- Jtr relative to what the pose tolerance allows it to move (b = Jt r vanishes at convergence, so a
difference of the two poses of d moves it by A d; 1e-5 is the pose tolerance); inlier counts within the borderline
correspondences a 1e-6 pose difference moves (exact at the first evaluation, see tests/test_ops_gpu.py)."""
import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

N_FRAMES = 31  # 30 tracked frames


def _render(w, h, n, seed=2024):
    import torch
    K = synth.Intrinsics.kinect(w, h)
    poses = synth.trajectory(n, seed=seed)
    dev, host = [], []
    for k in range(n):
        f = synth.render(poses[k], K, seed=seed, frame_id=k, device="cuda")
        dev.append(f)
        host.append({"depth": util.u16(f["depth"]), "rgba": f["rgba"].cpu().numpy(), "vmap": f["vmap"].cpu().numpy(),
                     "nmap": f["nmap"].cpu().numpy()})
    torch.cuda.synchronize()
    return K, poses.numpy(), dev, host


def _compare(tag, k, prod, t, R, tr, Rr, st, icp=True, rgb=True, so3=False):
    dt = float(np.abs(t - tr).max())
    dr = util.rot_err(R, Rr)
    assert dt <= 1e-5 and dr <= 1e-5, (tag, k, dt, dr)
    assert prod.se3_iterations == st["se3_iterations"], (tag, k)
    assert prod.so3_iterations == st["so3_iterations"], (tag, k)
    A, Ar, b, br = prod.lastA, st["last_A"], prod.lastb, st["last_b"]
    nA = float(np.linalg.norm(Ar))
    assert np.linalg.norm(A - Ar) <= 1e-4 * nA, (tag, k, np.linalg.norm(A - Ar) / nA)
    assert np.linalg.norm(b - br) <= 1e-4 * np.linalg.norm(br) + 1e-5 * np.linalg.norm(Ar, 2), (tag, k, np.linalg.norm(b - br), np.linalg.norm(br))
    if icp:
        assert abs(prod.lastICPCount - st["last_icp_count"]) <= 1e-4 * st["last_icp_count"], (tag, k, prod.lastICPCount, st["last_icp_count"])
        assert prod.lastICPError == pytest.approx(st["last_icp_error"], rel=1e-3), (tag, k)
    if rgb:
        assert abs(prod.lastRGBCount - st["last_rgb_count"]) <= max(2.0, 1e-4 * st["last_rgb_count"]), (tag, k, prod.lastRGBCount, st["last_rgb_count"])
    if so3:
        assert prod.lastSO3Count == st["last_so3_count"], (tag, k)


@pytest.mark.parametrize("cfg", [dict(size=(640, 480), so3=False), dict(size=(1280, 720), so3=True)], ids=["config1_640x480_joint", "config2_1280x720_so3"])
def test_trajectory_frame_by_frame(cfg):
    assert O.ref_available(), "oracle/_ref/libef_ref.so missing"
    w, h = cfg["size"]
    so3 = cfg["so3"]
    K, poses, dev, host = _render(w, h, N_FRAMES)
    posef = poses.astype(np.float32)
    m = dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=so3)
    kw = dict(rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=so3)
    single = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)  # ef_track_frame_to_model, device inputs
    five = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)    # the reference's five calls
    hostm = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_HOST)     # host-solve mode, five calls, host inputs
    ref = O.OracleTracker(w, h, K.cx, K.cy, K.fx, K.fy, impl="ref")
    worst = {"t": 0.0, "r": 0.0}
    try:
        if so3:  # once, before the first SO(3) call (ElasticFusion.cpp:326); the swap carries it from then on
            single.initFirstRGB(dev[0]["rgba"])
            five.initFirstRGB(dev[0]["rgba"])
            hostm.initFirstRGB(host[0]["rgba"])
            ref.init_first_rgb(host[0]["rgba"])
        for k in range(1, N_FRAMES):
            p = posef[k - 1]
            ref.init_icp_model(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
            ref.init_rgb_model(host[k - 1]["rgba"])
            ref.init_icp_depth(host[k]["depth"], 20.0)
            ref.init_rgb(host[k]["rgba"])
            tr, Rr, st = ref.get_incremental_transformation(p[:3, 3], p[:3, :3], **kw)

            t, R = single.trackFrameToModel(dev[k - 1]["vmap"], dev[k - 1]["nmap"], dev[k - 1]["rgba"], dev[k]["depth"], dev[k]["rgba"], 20.0, p,
                                            m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
            _compare("single-call", k, single, t, R, tr, Rr, st, so3=so3)
            worst["t"] = max(worst["t"], float(np.abs(t - tr).max()))
            worst["r"] = max(worst["r"], util.rot_err(R, Rr))

            five.initICPModel(dev[k - 1]["vmap"], dev[k - 1]["nmap"], 20.0, p)
            five.initRGBModel(dev[k - 1]["rgba"])
            five.initICP(dev[k]["depth"], 20.0)
            five.initRGB(dev[k]["rgba"])
            t5, R5 = five.getIncrementalTransformation(p[:3, 3], p[:3, :3], **m)
            _compare("five-call", k, five, t5, R5, tr, Rr, st, so3=so3)
            # the two entry points of the product agree to the bit, frame after frame
            assert np.array_equal(t5, t) and np.array_equal(R5, R), k

            hostm.initICPModel(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
            hostm.initRGBModel(host[k - 1]["rgba"])
            hostm.initICP(host[k]["depth"], 20.0)
            hostm.initRGB(host[k]["rgba"])
            th, Rh = hostm.getIncrementalTransformation(p[:3, 3], p[:3, :3], **m)
            _compare("host-solve", k, hostm, th, Rh, tr, Rr, st, so3=so3)

            # and the tracker tracks: closer to the ground truth of frame k than the prior was
            gt = poses[k]
            assert np.linalg.norm(t - gt[:3, 3]) < np.linalg.norm(p[:3, 3] - gt[:3, 3]) + 1e-4, k
        print(f"\n[trajectory parity {w}x{h} so3={so3}] worst |dt| {worst['t']:.2e} m, worst rotation {worst['r']:.2e} rad over {N_FRAMES - 1} frames")
    finally:
        for x in (single, five, hostm):
            x.close()
        ref.close()
