"""Real-data parity: the reference's only shipped fixture, elasticfusionpublic/GPUTest/{1c,1d,2c,2d}.png (a Kinect frame
pair), turned into tracker inputs exactly as GPUTest.cpp does (tests/util.py::gputest_inputs, fixture decoded by
tests/golden/make_gputest_fixture.py) and tracked in every mode by the product and by the reference's CUDA kernels."""
import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

MODES = {
    "gputest": dict(rgbOnly=False, icpWeight=10.0, pyramid=False, fastOdom=False, so3=True),  # the call GPUTest.cpp:278 makes
    "icp_only": dict(rgbOnly=False, icpWeight=100.0, pyramid=True, fastOdom=False, so3=False),
    "joint": dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=False),
    "joint_so3": dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True),
    "rgb_only": dict(rgbOnly=True, icpWeight=10.0, pyramid=True, fastOdom=False, so3=False),
    "fast_nopyr": dict(rgbOnly=False, icpWeight=10.0, pyramid=False, fastOdom=True, so3=True),
}


def _kw(m):
    return dict(rgb_only=m["rgbOnly"], icp_weight=m["icpWeight"], pyramid=m["pyramid"], fast_odom=m["fastOdom"], so3=m["so3"])


def _feed(tr, pose, f0, f1):
    if isinstance(tr, O.OracleTracker):
        tr.init_first_rgb(f0["rgba"])
        tr.init_icp_model(f0["vmap"], f0["nmap"], 20.0, pose)
        tr.init_rgb_model(f0["rgba"])
        tr.init_icp_depth(f1["depth"], 20.0)
        tr.init_rgb(f1["rgba"])
    else:
        tr.initFirstRGB(f0["rgba"])
        tr.initICPModel(f0["vmap"], f0["nmap"], 20.0, pose)
        tr.initRGBModel(f0["rgba"])
        tr.initICP(f1["depth"], 20.0)
        tr.initRGB(f1["rgba"])


@pytest.mark.parametrize("solve_mode", [RO.EF_SOLVE_HOST, RO.EF_SOLVE_DEVICE])
def test_gputest_png_pair_matches_reference_cuda(solve_mode):
    assert O.ref_available(), "oracle/_ref/libef_ref.so missing"
    K, f0, f1 = util.gputest_inputs()
    w, h = 640, 480
    pose = np.eye(4, dtype=np.float32)  # GPUTest.cpp:213 currPose = Identity
    prod = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=solve_mode)
    ens = util.RefEnsemble(w, h, K)  # the reference at GPUConfig's defaults + at other launch shapes (its own spread)
    ref = ens.ref
    try:
        _feed(prod, pose, f0, f1)
        _feed(ref, pose, f0, f1)
        for lvl in range(3):
            rows = h >> lvl
            for name in ("depth_tmp", "last_image", "next_image", "last_next_image"):
                assert np.array_equal(prod.buffer(name, lvl), ref.buffer(name, lvl)), (name, lvl)
            for name in ("vmap_curr", "nmap_curr", "vmap_g_prev", "nmap_g_prev"):
                assert util.masked_map_compare(prod.buffer(name, lvl), ref.buffer(name, lvl), rows) <= 1, (name, lvl)
            for name in ("last_depth", "next_depth"):
                assert int(util.ulp_diff(prod.buffer(name, lvl), ref.buffer(name, lvl)).max()) <= 1, (name, lvl)
        for name, m in MODES.items():
            _feed(prod, pose, f0, f1)
            ens.each(lambda r: _feed(r, pose, f0, f1))
            t, R = prod.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **m)
            tr, Rr, st, spread = ens.track(pose[:3, 3], pose[:3, :3], **_kw(m))
            dt, dr = float(np.abs(t - tr).max()), util.rot_err(R, Rr)
            # BASELINE's 1e-5, or K times what the reference moves against itself on this real, noisy pair when only its launch
            # shape changes (float sums in a different order; measured 1e-5 .. 3e-5 in the no-pyramid and RGB-only modes)
            assert dt <= max(1e-5, util.RefEnsemble.K * spread["t"]) and dr <= max(1e-5, util.RefEnsemble.K * spread["r"]), (name, dt, dr, spread)
            print(f"\n[gputest pair {name}] |dt| {dt:.1e} m, rotation {dr:.1e} rad (reference vs itself: {spread['t']:.1e} / {spread['r']:.1e})")
            assert prod.se3_iterations == st["se3_iterations"] and prod.so3_iterations == st["so3_iterations"], name
            if not m["rgbOnly"]:
                assert abs(prod.lastICPCount - st["last_icp_count"]) <= max(2e-4 * st["last_icp_count"], util.RefEnsemble.K * spread["icp"]), (name, spread)
            if m["so3"]:
                assert prod.lastSO3Count == st["last_so3_count"], name
            if m["rgbOnly"] or m["icpWeight"] < 100:
                for lvl in range(3):
                    assert np.array_equal(prod.buffer("dIdx", lvl), ref.buffer("dIdx", lvl))
                    assert np.array_equal(prod.buffer("dIdy", lvl), ref.buffer("dIdy", lvl))
            Ar = st["last_A"]
            assert np.linalg.norm(prod.lastA - Ar) <= max(1e-4, util.RefEnsemble.K * spread["A"]) * np.linalg.norm(Ar), (name, np.linalg.norm(prod.lastA - Ar) / np.linalg.norm(Ar), spread)
            # the frames are ~3 cm / ~1 degree apart: the tracker must have moved
            if name in ("joint", "icp_only", "joint_so3"):
                assert 0.003 < float(np.linalg.norm(t)) < 0.2, (name, t)
    finally:
        prod.close()
        ens.close()


def test_gputest_pair_single_call_and_arrays():
    """the same pair through ef_track_frame_to_model (device inputs) and through the cudaArray entry points: same bits"""
    import torch
    K, f0, f1 = util.gputest_inputs()
    w, h = 640, 480
    pose = np.eye(4, dtype=np.float32)
    m = MODES["joint"]
    a = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    b = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    c = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    arrs = []
    try:
        _feed(a, pose, f0, f1)
        ta, Ra = a.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **m)
        dev = [torch.from_numpy(x).cuda() for x in (f0["vmap"], f0["nmap"], f0["rgba"])]
        dev.append(torch.from_numpy(f1["depth"].view(np.int16)).cuda().view(torch.uint16))
        dev.append(torch.from_numpy(f1["rgba"]).cuda())
        tb, Rb = b.trackFrameToModel(*dev, 20.0, pose, m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
        assert np.array_equal(ta, tb) and np.array_equal(Ra, Rb)
        arrs = [util.CudaArray(f0["vmap"], "rgba32f"), util.CudaArray(f0["nmap"], "rgba32f"), util.CudaArray(f0["rgba"], "rgba8"),
                util.CudaArray(f1["depth"], "u16"), util.CudaArray(f1["rgba"], "rgba8")]
        c.initICPModel(arrs[0], arrs[1], 20.0, pose)
        c.initRGBModel(arrs[2])
        c.initICP(arrs[3], 20.0)
        c.initRGB(arrs[4])
        tc, Rc = c.getIncrementalTransformation(pose[:3, 3], pose[:3, :3], **m)
        assert np.array_equal(ta, tc) and np.array_equal(Ra, Rc)
    finally:
        for x in (a, b, c):
            x.close()
        for x in arrs:
            x.free()
