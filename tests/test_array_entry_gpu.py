"""The six cudaArray entry points (ef_init_*_array): what a CUDA<->GL interop caller -- the RGBDOdometry shim of
include/compat/RGBDOdometry.h over GPUTexture::cudaRes, reference RGBDOdometry.cpp:120-128 -- feeds the tracker with.
Every pyramid must come out bit-identical to the pointer / host entry points, with and without EF_OPT_DEFER_BUILD, and
back-to-back calls that reuse the staging buffers must not race (round-1 advisor finding)."""
import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from tests import util

pytestmark = pytest.mark.gpu

JOINT_SO3 = dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True)
NAMES_EXACT = ("depth_tmp", "last_image", "next_image", "last_next_image", "vmap_curr", "nmap_curr", "vmap_g_prev", "nmap_g_prev",
               "last_depth", "next_depth")


def _arrays(f0, f1):
    return {"v0": util.CudaArray(f0["vmap"], "rgba32f"), "n0": util.CudaArray(f0["nmap"], "rgba32f"),
            "c0": util.CudaArray(f0["rgba"], "rgba8"), "d1": util.CudaArray(f1["depth"], "u16"), "c1": util.CudaArray(f1["rgba"], "rgba8"),
            "v1": util.CudaArray(f1["vmap"], "rgba32f"), "n1": util.CudaArray(f1["nmap"], "rgba32f")}


def _same_pyramids(a, b, h):
    for lvl in range(3):
        for name in NAMES_EXACT:
            assert np.array_equal(a.buffer(name, lvl), b.buffer(name, lvl), equal_nan=True), (name, lvl)


@pytest.mark.parametrize("defer", [0, 1])
@pytest.mark.parametrize("solve_mode", [RO.EF_SOLVE_HOST, RO.EF_SOLVE_DEVICE])
@pytest.mark.parametrize("size", [(640, 480), (320, 240)])
def test_array_entry_points_equal_the_host_entry_points(size, solve_mode, defer):
    w, h = size
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    f1 = dict(f1, depth=util.punch_holes(f1["depth"]))
    pose0f = pose0.astype(np.float32)
    A = _arrays(f0, f1)
    a = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=solve_mode)  # cudaArray inputs
    b = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=solve_mode)  # host inputs
    a.set_option(RO.EF_OPT_DEFER_BUILD, defer)
    try:
        for rep in range(3):  # repeated frames: the staging buffers are reused while builders of the previous call may still run
            a.initFirstRGB(A["c0"])
            a.initICPModel(A["v0"], A["n0"], 20.0, pose0f)
            a.initRGBModel(A["c0"])
            a.initICP(A["d1"], 20.0)
            a.initRGB(A["c1"])  # must not overwrite what initRGBModel staged (one buffer each)
            b.initFirstRGB(f0["rgba"])
            b.initICPModel(f0["vmap"], f0["nmap"], 20.0, pose0f)
            b.initRGBModel(f0["rgba"])
            b.initICP(f1["depth"], 20.0)
            b.initRGB(f1["rgba"])
            if rep == 0:
                _same_pyramids(a, b, h)
                a.initFirstRGB(A["c0"])  # the download flushed the deferred build and synchronised: feed again, undisturbed
                a.initICPModel(A["v0"], A["n0"], 20.0, pose0f)
                a.initRGBModel(A["c0"])
                a.initICP(A["d1"], 20.0)
                a.initRGB(A["c1"])
            ta, Ra = a.getIncrementalTransformation(pose0f[:3, 3], pose0f[:3, :3], **JOINT_SO3)
            tb, Rb = b.getIncrementalTransformation(pose0f[:3, 3], pose0f[:3, :3], **JOINT_SO3)
            assert np.array_equal(ta, tb) and np.array_equal(Ra, Rb), (rep, ta, tb)
            assert a.lastICPCount == b.lastICPCount and a.lastRGBCount == b.lastRGBCount and a.lastSO3Count == b.lastSO3Count
            assert np.array_equal(a.lastA, b.lastA) and np.array_equal(a.lastb, b.lastb)
        if defer:
            # all four inputs recorded -> one builder launch + the tracker kernel (device mode); never more launches than undeferred
            l0 = a.launch_count
            a.initICPModel(A["v0"], A["n0"], 20.0, pose0f)
            a.initRGBModel(A["c0"])
            a.initICP(A["d1"], 20.0)
            a.initRGB(A["c1"])
            assert a.launch_count == l0  # nothing built yet
            a.getIncrementalTransformation(pose0f[:3, 3], pose0f[:3, :3], **dict(JOINT_SO3, so3=False))
            if solve_mode == RO.EF_SOLVE_DEVICE:
                assert a.launch_count - l0 == 2
        # the maps overload (modelToModel / Ferns: initICP(vertices, normals)) through cudaArrays
        a.initICP(A["v1"], A["n1"], 20.0)
        b.initICP(f1["vmap"], f1["nmap"], 20.0)
        for lvl in range(3):
            for name in ("vmap_curr", "nmap_curr"):
                assert np.array_equal(a.buffer(name, lvl), b.buffer(name, lvl), equal_nan=True), (name, lvl)
    finally:
        a.close()
        b.close()
        for x in A.values():
            x.free()


def test_same_entry_twice_keeps_both_calls_apart():
    """initRGBModel twice in a row, then initRGB: each call's pyramid comes from its own image even when the builder of the
    first call may still be reading the staging buffer the second call overwrites (stage_guard in ef_api.cu)."""
    w, h = 640, 480
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    pose0f = pose0.astype(np.float32)
    black = np.zeros_like(f0["rgba"])
    A = _arrays(f0, f1)
    Ab = util.CudaArray(black, "rgba8")
    for defer in (0, 1):
        a = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
        b = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
        a.set_option(RO.EF_OPT_DEFER_BUILD, defer)
        try:
            a.initICPModel(A["v0"], A["n0"], 20.0, pose0f)
            a.initRGBModel(Ab)
            a.initRGBModel(A["c0"])
            a.initRGB(A["c1"])
            b.initICPModel(f0["vmap"], f0["nmap"], 20.0, pose0f)
            b.initRGBModel(f0["rgba"])
            b.initRGB(f1["rgba"])
            for lvl in range(3):
                assert np.array_equal(a.buffer("last_image", lvl), b.buffer("last_image", lvl)), (defer, lvl)
                assert np.array_equal(a.buffer("next_image", lvl), b.buffer("next_image", lvl)), (defer, lvl)
        finally:
            a.close()
            b.close()
    Ab.free()
    for x in A.values():
        x.free()
