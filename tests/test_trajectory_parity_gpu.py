"""Frame-by-frame parity over consecutive frames of BASELINE.json's trajectories (configs[1] and configs[2]) against the
reference's own CUDA kernels under the restated host loop (oracle/_ref), with the tracker state CARRIED across frames
(the SO(3) pre-alignment swaps its image pyramids at the end of every call, RGBDOdometry.cpp:593-599 -- nothing is re-fed).

Open-loop protocol of SURVEY.md 8(d): model maps of frame k are ray-cast at the ground-truth pose k-1 and the tracker is
called with that pose as the prior.  Tolerances (BASELINE.json north_star): pose 1e-5 m / 1e-5 rad per frame; JtJ | Jtr
1e-4 norm-relative -- Jtr relative to what the pose tolerance allows it to move (b = Jt r vanishes at convergence, so a
difference of the two poses of d moves it by A d); inlier counts within the borderline correspondences a 1e-6 pose
difference moves (exact at the first evaluation, see tests/test_ops_gpu.py).  On a few frames of either trajectory the
reference does not reproduce ITSELF to these tolerances when only its launch shape changes (util.RefEnsemble,
profiles/r02_parity_spread.txt: up to 4e-4 m on frame 1 of config 1); there the bound is K times the reference's own
spread on that frame, and the test also asserts that every frame outside 1e-5 is such a frame."""
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


def _compare(tag, k, prod, t, R, tr, Rr, st, spread, devs, so3=False):
    """tolerance = BASELINE's, or the reference's own launch-shape spread on this frame where that is larger (util.RefEnsemble)"""
    dt = float(np.abs(t - tr).max())
    dr = util.rot_err(R, Rr)
    devs.append((dt, dr))
    assert dt <= max(1e-5, util.RefEnsemble.K * spread["t"]) and dr <= max(1e-5, util.RefEnsemble.K * spread["r"]), (tag, k, dt, dr, spread)
    assert prod.se3_iterations == st["se3_iterations"], (tag, k)
    assert prod.so3_iterations == st["so3_iterations"], (tag, k)
    A, Ar, b, br = prod.lastA, st["last_A"], prod.lastb, st["last_b"]
    nA = float(np.linalg.norm(Ar))
    assert np.linalg.norm(A - Ar) <= max(1e-4, util.RefEnsemble.K * spread["A"]) * nA, (tag, k, np.linalg.norm(A - Ar) / nA, spread)
    # b = Jt r vanishes at convergence: a pose difference d between two evaluations moves it by A d
    tol_b = 1e-4 * np.linalg.norm(br) + max(1e-5, util.RefEnsemble.K * spread["t"], util.RefEnsemble.K * spread["r"]) * np.linalg.norm(Ar, 2) + util.RefEnsemble.K * spread["b"]
    assert np.linalg.norm(b - br) <= tol_b, (tag, k, np.linalg.norm(b - br), tol_b)
    assert abs(prod.lastICPCount - st["last_icp_count"]) <= max(2.0, util.RefEnsemble.K * spread["icp"], 1e-5 * st["last_icp_count"]), (tag, k, prod.lastICPCount, st["last_icp_count"], spread)
    assert prod.lastICPError == pytest.approx(st["last_icp_error"], rel=max(1e-3, util.RefEnsemble.K * spread["icp_err"])), (tag, k)
    assert abs(prod.lastRGBCount - st["last_rgb_count"]) <= max(2.0, util.RefEnsemble.K * spread["rgb"], 1e-5 * st["last_rgb_count"]), (tag, k, prod.lastRGBCount, st["last_rgb_count"], spread)
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
    ens = util.RefEnsemble(w, h, K)
    devs = {"single-call": [], "five-call": [], "host-solve": []}
    spreads = []
    try:
        if so3:  # once, before the first SO(3) call (ElasticFusion.cpp:326); the swap carries it from then on
            single.initFirstRGB(dev[0]["rgba"])
            five.initFirstRGB(dev[0]["rgba"])
            hostm.initFirstRGB(host[0]["rgba"])
            ens.each(lambda r: r.init_first_rgb(host[0]["rgba"]))
        for k in range(1, N_FRAMES):
            p = posef[k - 1]

            def feed(r):
                r.init_icp_model(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
                r.init_rgb_model(host[k - 1]["rgba"])
                r.init_icp_depth(host[k]["depth"], 20.0)
                r.init_rgb(host[k]["rgba"])
            ens.each(feed)
            tr, Rr, st, spread = ens.track(p[:3, 3], p[:3, :3], **kw)
            spreads.append((spread["t"], spread["r"]))

            t, R = single.trackFrameToModel(dev[k - 1]["vmap"], dev[k - 1]["nmap"], dev[k - 1]["rgba"], dev[k]["depth"], dev[k]["rgba"], 20.0, p,
                                            m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
            _compare("single-call", k, single, t, R, tr, Rr, st, spread, devs["single-call"], so3=so3)

            five.initICPModel(dev[k - 1]["vmap"], dev[k - 1]["nmap"], 20.0, p)
            five.initRGBModel(dev[k - 1]["rgba"])
            five.initICP(dev[k]["depth"], 20.0)
            five.initRGB(dev[k]["rgba"])
            t5, R5 = five.getIncrementalTransformation(p[:3, 3], p[:3, :3], **m)
            _compare("five-call", k, five, t5, R5, tr, Rr, st, spread, devs["five-call"], so3=so3)
            # the two entry points of the product agree to the bit, frame after frame
            assert np.array_equal(t5, t) and np.array_equal(R5, R), k

            hostm.initICPModel(host[k - 1]["vmap"], host[k - 1]["nmap"], 20.0, p)
            hostm.initRGBModel(host[k - 1]["rgba"])
            hostm.initICP(host[k]["depth"], 20.0)
            hostm.initRGB(host[k]["rgba"])
            th, Rh = hostm.getIncrementalTransformation(p[:3, 3], p[:3, :3], **m)
            _compare("host-solve", k, hostm, th, Rh, tr, Rr, st, spread, devs["host-solve"], so3=so3)

            # and the tracker tracks: closer to the ground truth of frame k than the prior was
            gt = poses[k]
            assert np.linalg.norm(t - gt[:3, 3]) < np.linalg.norm(p[:3, 3] - gt[:3, 3]) + 1e-4, k
        # over the whole trajectory: the typical frame is far inside BASELINE's tolerance, and the frames outside it are the
        # ones on which the reference does not reproduce itself either
        sp = np.array(spreads)
        for tag, d in devs.items():
            d = np.array(d)
            assert np.median(d[:, 0]) <= 2e-6 and np.median(d[:, 1]) <= 2e-6, (tag, np.median(d, axis=0))
            inside = (d[:, 0] <= 1e-5) & (d[:, 1] <= 1e-5)
            assert inside.mean() >= 0.85, (tag, inside.mean())
            assert np.all(inside | (sp[:, 0] > 5e-6) | (sp[:, 1] > 5e-6)), (tag, d[~inside], sp[~inside])
            print(f"\n[trajectory parity {w}x{h} so3={so3} {tag}] median |dt| {np.median(d[:, 0]):.1e} m / {np.median(d[:, 1]):.1e} rad, "
                  f"{int(inside.sum())}/{len(d)} frames within 1e-5, worst {d[:, 0].max():.1e} m / {d[:, 1].max():.1e} rad "
                  f"(reference vs itself at other launch shapes: worst {sp[:, 0].max():.1e} m / {sp[:, 1].max():.1e} rad)")
    finally:
        for x in (single, five, hostm):
            x.close()
        ens.close()
