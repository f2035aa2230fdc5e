"""The reference-side binding as CODE: include/compat/RGBDOdometry.h (class RGBDOdometry over the C ABI) and
include/compat/cudafuncs.cuh (the 17 operator functions of Cuda/cudafuncs.cuh:64-177 over the caller's DeviceArray2D),
compiled by tests/compat/compat_harness.cu against the reference's own containers and types and run here next to the
reference's CUDA kernels (oracle/_ref).  The harness library is test infrastructure (oracle/_ref/libef_compat_test.so)."""
import ctypes as C
import os

import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth
from oracle import oracle as O
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libef_compat_test.so")
HEADERS = [os.path.join(ROOT, "include", "compat", n) for n in ("RGBDOdometry.h", "cudafuncs.cuh")]


def test_binding_headers_exist_and_name_every_reference_function():
    """CPU-side: the two headers are there and declare the reference's interface (names as in cudafuncs.cuh / RGBDOdometry.h)."""
    shim = open(HEADERS[0]).read()
    for name in ("class RGBDOdometry", "void initICP(GPUTexture * filteredDepth", "void initICP(GPUTexture * predictedVertices",
                 "void initICPModel(", "void initRGB(", "void initRGBModel(", "void initFirstRGB(", "void getIncrementalTransformation(",
                 "Eigen::MatrixXd getCovariance()", "lastICPError", "lastICPCount", "lastRGBError", "lastRGBCount", "lastSO3Error", "lastSO3Count",
                 "lastA", "lastb", "float distThresh = 0.10f", "sin(20.f * 3.14159254f / 180.f)"):
        assert name in shim, name
    ops = open(HEADERS[1]).read()
    for fn in ("icpStep", "rgbStep", "so3Step", "computeRgbResidual", "createVMap", "createNMap", "tranformMaps", "copyMaps", "resizeVMap",
               "resizeNMap", "imageBGRToIntensity", "verticesToDepth", "projectToPointCloud", "pyrDown", "pyrDownGaussF", "pyrDownUcharGauss",
               "computeDerivativeImages"):
        assert f"inline void {fn}(" in ops, fn


def _lib():
    assert os.path.exists(SO), "oracle/_ref/libef_compat_test.so missing: make -C oracle ref (where /root/reference exists)"
    L = C.CDLL(SO)
    L.efc_shim_create.restype = C.c_void_p
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.gpu
def test_shim_class_tracks_like_the_reference_and_like_the_python_mirror():
    """five consecutive frames of config 1 with SO(3) pre-alignment (state carried across frames) through
    `class RGBDOdometry` of include/compat/RGBDOdometry.h"""
    L = _lib()
    w, h = 640, 480
    K = synth.Intrinsics.kinect(w, h)
    n = 6
    poses = synth.trajectory(n, seed=2024)
    frames = []
    for k in range(n):
        f = synth.render(poses[k], K, seed=2024, frame_id=k, device="cuda")
        frames.append({"depth": util.u16(f["depth"]), "rgba": f["rgba"].cpu().numpy(), "vmap": f["vmap"].cpu().numpy(), "nmap": f["nmap"].cpu().numpy()})
    posef = poses.numpy().astype(np.float32)
    shim = C.c_void_p(L.efc_shim_create(w, h, C.c_float(K.cx), C.c_float(K.cy), C.c_float(K.fx), C.c_float(K.fy)))
    assert shim.value
    mirror = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    ens = util.RefEnsemble(w, h, K)
    err = C.create_string_buffer(256)
    try:
        mirror.initFirstRGB(frames[0]["rgba"])
        ens.each(lambda r: r.init_first_rgb(frames[0]["rgba"]))
        for k in range(1, n):
            p = np.ascontiguousarray(posef[k - 1])
            f0, f1 = frames[k - 1], frames[k]
            t, R, st6 = np.zeros(3, np.float32), np.zeros(9, np.float32), np.zeros(6, np.float32)
            A, b, cov = np.zeros(36), np.zeros(6), np.zeros(36)
            rc = L.efc_shim_frame(shim, int(k == 1), _p(f0["vmap"]), _p(f0["nmap"]), _p(f0["rgba"]), _p(f1["depth"]), _p(f1["rgba"]), _p(p), 0,
                                  C.c_float(10.0), 1, 0, 1, _p(t), _p(R), _p(st6), _p(A), _p(b), _p(cov), err, 256)
            assert rc == 0, err.value
            # == the Python mirror's five calls (same library underneath): bit for bit
            mirror.initICPModel(f0["vmap"], f0["nmap"], 20.0, p)
            mirror.initRGBModel(f0["rgba"])
            mirror.initICP(f1["depth"], 20.0)
            mirror.initRGB(f1["rgba"])
            tm, Rm = mirror.getIncrementalTransformation(p[:3, 3], p[:3, :3], False, 10.0, True, False, True)
            assert np.array_equal(t, tm) and np.array_equal(R.reshape(3, 3), Rm), k
            assert np.array_equal(A.reshape(6, 6), mirror.lastA) and np.array_equal(b, mirror.lastb)
            assert st6[1] == mirror.lastICPCount and st6[3] == mirror.lastRGBCount and st6[5] == mirror.lastSO3Count
            assert np.allclose(cov.reshape(6, 6) @ A.reshape(6, 6), np.eye(6), atol=1e-6)

            # vs the reference's CUDA tracker
            def feed(r):
                r.init_icp_model(f0["vmap"], f0["nmap"], 20.0, p)
                r.init_rgb_model(f0["rgba"])
                r.init_icp_depth(f1["depth"], 20.0)
                r.init_rgb(f1["rgba"])
            ens.each(feed)
            tr, Rr, st, spread = ens.track(p[:3, 3], p[:3, :3], rgb_only=False, icp_weight=10.0, pyramid=True, fast_odom=False, so3=True)
            dt, dr = float(np.abs(t - tr).max()), util.rot_err(R.reshape(3, 3), Rr)
            assert dt <= max(1e-5, util.RefEnsemble.K * spread["t"]) and dr <= max(1e-5, util.RefEnsemble.K * spread["r"]), (k, dt, dr, spread)
            assert st6[5] == st["last_so3_count"]
    finally:
        L.efc_shim_destroy(shim)
        mirror.close()
        ens.close()


@pytest.mark.gpu
def test_shim_class_drop_in_rate_in_cxx():
    """The drop-in number measured where a drop-in lives: a C++ loop over `class RGBDOdometry` of include/compat/RGBDOdometry.h --
    textures resident as cudaArrays (the reference's GL textures), the five init calls + getIncrementalTransformation per frame, the
    pose read back every frame.  (bench.py's value_reference_api drives the same five C-ABI calls from Python.)  The rate is recorded
    in gpurun_out/compat_shim_fps.txt; the assertion is only a sanity floor (the reference's own kernels reach ~350 frames/s)."""
    L = _lib()
    L.efc_shim_bench.restype = C.c_double
    w, h = 640, 480
    K = synth.Intrinsics.kinect(w, h)
    poses = synth.trajectory(2, seed=2024)
    f0 = synth.render(poses[0], K, seed=2024, frame_id=0, device="cuda")
    f1 = synth.render(poses[1], K, seed=2024, frame_id=1, device="cuda")
    p = np.ascontiguousarray(poses.numpy().astype(np.float32)[0])
    v, n, m = f0["vmap"].cpu().numpy(), f0["nmap"].cpu().numpy(), f0["rgba"].cpu().numpy()
    d, c = util.u16(f1["depth"]), f1["rgba"].cpu().numpy()
    shim = C.c_void_p(L.efc_shim_create(w, h, C.c_float(K.cx), C.c_float(K.cy), C.c_float(K.fx), C.c_float(K.fy)))
    assert shim.value
    try:
        rates = {}
        for name, defer in (("five calls", 0), ("five calls, EF_OPT_DEFER_BUILD", 1)):
            rates[name] = float(L.efc_shim_bench(shim, _p(v), _p(n), _p(m), _p(d), _p(c), _p(p), 300, defer))
            assert rates[name] > 1000.0, rates
        out = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "compat_shim_fps.txt"), "w") as fh:
                for name, r in rates.items():
                    fh.write(f"class RGBDOdometry (include/compat/RGBDOdometry.h), C++ loop, 640x480 joint ICP+RGB, {name}: {r:.0f} frames/s\n")
    finally:
        L.efc_shim_destroy(shim)


@pytest.mark.gpu
def test_operator_functions_over_the_reference_containers():
    """pyrDown / createVMap / createNMap / imageBGRToIntensity / pyrDownUcharGauss / computeDerivativeImages / copyMaps /
    resizeVMap / resizeNMap / tranformMaps / icpStep through include/compat/cudafuncs.cuh on pitched DeviceArray2D buffers
    (cudaMallocPitch: 512-byte aligned rows) against the reference kernels"""
    L = _lib()
    w, h = 640, 480
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    depth = util.punch_holes(f1["depth"])
    err = C.create_string_buffer(256)
    r2, c2 = h // 2, w // 2

    d1 = np.zeros((r2, c2), np.uint16)
    v1 = np.zeros((3 * r2, c2), np.float32)
    n1 = np.zeros((3 * r2, c2), np.float32)
    rc = L.efc_ops_depth_chain(_p(depth), h, w, C.c_float(K.fx), C.c_float(K.fy), C.c_float(K.cx), C.c_float(K.cy), C.c_float(20.0), _p(d1), _p(v1),
                               _p(n1), err, 256)
    assert rc == 0, err.value
    d1r = O.pyr_down_u16(depth, impl="ref")
    assert np.array_equal(d1, d1r)
    fx, fy, cx, cy = util.se3_level_params(K, 1)
    v1r = O.create_vmap(d1r, fx, fy, cx, cy, 20.0, impl="ref")
    n1r = O.create_nmap(v1r, impl="ref")
    assert util.masked_map_compare(v1, v1r, r2) == 0
    assert util.masked_map_compare(n1, n1r, r2) <= 1

    i0, i1 = np.zeros((h, w), np.uint8), np.zeros((r2, c2), np.uint8)
    gx, gy = np.zeros((r2, c2), np.int16), np.zeros((r2, c2), np.int16)
    rc = L.efc_ops_image_chain(_p(f1["rgba"]), h, w, _p(i0), _p(i1), _p(gx), _p(gy), err, 256)
    assert rc == 0, err.value
    i0r = O.bgr_to_intensity(f1["rgba"], impl="ref")
    i1r = O.pyr_down_gauss_u8(i0r, impl="ref")
    gxr, gyr = O.derivative_images(i1r, impl="ref")
    assert np.array_equal(i0, i0r) and np.array_equal(i1, i1r) and np.array_equal(gx, gxr) and np.array_equal(gy, gyr)

    pose = np.ascontiguousarray(pose0.astype(np.float32))
    A, b, res = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
    rc = L.efc_ops_icp(_p(f0["vmap"]), _p(f0["nmap"]), _p(depth), h, w, C.c_float(K.fx), C.c_float(K.fy), C.c_float(K.cx), C.c_float(K.cy), _p(pose),
                       _p(A), _p(b), _p(res), err, 256)
    assert rc == 0, err.value
    # the same chain through the reference kernels
    vp0, np0 = O.copy_maps(f0["vmap"], f0["nmap"], impl="ref")
    vp1, np1 = O.resize_map(vp0, False, impl="ref"), O.resize_map(np0, True, impl="ref")
    R, t = pose[:3, :3].copy(), pose[:3, 3].copy()
    vp1, np1 = O.transform_maps(vp1, np1, R, t, impl="ref")
    Rinv = np.ascontiguousarray(R.T)
    ang = float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0)))
    Ar, br, resr = O.icp_step(R, t, v1r, n1r, Rinv, t, fx, fy, cx, cy, vp1, np1, 0.10, ang, impl="ref")
    assert res[1] == resr[1] and res[1] > 1000
    assert np.linalg.norm(A.reshape(6, 6) - Ar) <= 1e-4 * np.linalg.norm(Ar)
    assert np.linalg.norm(b - br) <= 1e-4 * np.linalg.norm(br)
