"""Tier-2 per-operator entry points (ef_op_* in include/ef_track.h) with the names of the reference's
free functions (elasticfusionpublic/Core/src/Cuda/cudafuncs.cuh:64-177).

Inputs are numpy arrays or torch tensors; they are placed on the current CUDA device with torch
(device memory plumbing only), the operator runs through the C ABI on torch's current stream, and the
result comes back as numpy.  Dense rows (pitch 0) unless `pitch_bytes` is given by the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import binding
from .binding import EFError

DATA_TERM = np.dtype([("zero_x", np.int16), ("zero_y", np.int16), ("one_x", np.int16), ("one_y", np.int16),
                      ("diff", np.float32), ("valid", np.uint8), ("pad", np.uint8, (3,))])


def _dev(a, dtype):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.ascontiguousarray(a)
        if arr.dtype == np.uint16:
            t = torch.from_numpy(arr.view(np.int16)).view(torch.uint16)
        else:
            t = torch.from_numpy(arr)
    t = t.to("cuda").contiguous()
    assert t.dtype == dtype, (t.dtype, dtype)
    return t


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc, where):
    if rc != 0:
        raise EFError(rc, where)


def _np(t):
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _f9(m):
    return np.ascontiguousarray(np.asarray(m, np.float32).reshape(-1))


def _scratch():
    n = binding.lib().ef_op_scratch_bytes()
    return torch.zeros(n, dtype=torch.uint8, device="cuda")


def pyrDown(src):
    L = binding.lib()
    s = _dev(src, torch.uint16)
    r, c = s.shape
    d = torch.zeros((r // 2, c // 2), dtype=torch.uint16, device="cuda")
    _chk(L.ef_op_pyr_down_u16(_p(s), C.c_size_t(0), r, c, _p(d), C.c_size_t(0), _stream()), "ef_op_pyr_down_u16")
    return _np(d)


def createVMap(depth, fx, fy, cx, cy, cutoff):
    L = binding.lib()
    d = _dev(depth, torch.uint16)
    r, c = d.shape
    v = torch.zeros((3 * r, c), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_create_vmap(_p(d), C.c_size_t(0), r, c, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                             C.c_float(cutoff), _p(v), C.c_size_t(0), _stream()), "ef_op_create_vmap")
    return _np(v)


def createNMap(vmap):
    L = binding.lib()
    v = _dev(vmap, torch.float32)
    r, c = v.shape[0] // 3, v.shape[1]
    n = torch.zeros_like(v)
    _chk(L.ef_op_create_nmap(_p(v), C.c_size_t(0), r, c, _p(n), C.c_size_t(0), _stream()), "ef_op_create_nmap")
    return _np(n)


def tranformMaps(vmap, nmap, R, t):
    L = binding.lib()
    v = _dev(vmap, torch.float32).clone()
    n = _dev(nmap, torch.float32).clone()
    r, c = v.shape[0] // 3, v.shape[1]
    Rm, tv = _f9(R), _f9(t)
    _chk(L.ef_op_transform_maps(_p(v), _p(n), C.c_size_t(0), r, c, Rm.ctypes.data_as(C.c_void_p), tv.ctypes.data_as(C.c_void_p),
                                _p(v), _p(n), C.c_size_t(0), _stream()), "ef_op_transform_maps")
    return _np(v), _np(n)


def copyMaps(v4, n4):
    L = binding.lib()
    a, b = _dev(v4, torch.float32), _dev(n4, torch.float32)
    r, c = a.shape[:2]
    v = torch.zeros((3 * r, c), dtype=torch.float32, device="cuda")
    n = torch.zeros((3 * r, c), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_copy_maps(_p(a), _p(b), r, c, _p(v), _p(n), C.c_size_t(0), _stream()), "ef_op_copy_maps")
    return _np(v), _np(n)


def _resize(m, normalize):
    L = binding.lib()
    i = _dev(m, torch.float32)
    r, c = i.shape[0] // 3, i.shape[1]
    o = torch.zeros((3 * (r // 2), c // 2), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_resize_map(_p(i), C.c_size_t(0), r, c, _p(o), C.c_size_t(0), int(normalize), _stream()), "ef_op_resize_map")
    return _np(o)


def resizeVMap(m):
    return _resize(m, False)


def resizeNMap(m):
    return _resize(m, True)


def verticesToDepth(v4, cutoff):
    L = binding.lib()
    a = _dev(v4, torch.float32)
    r, c = a.shape[:2]
    d = torch.zeros((r, c), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_vertices_to_depth(_p(a), r, c, C.c_float(cutoff), _p(d), C.c_size_t(0), _stream()), "ef_op_vertices_to_depth")
    return _np(d)


def pyrDownGaussF(src):
    L = binding.lib()
    s = _dev(src, torch.float32)
    r, c = s.shape
    d = torch.zeros((r // 2, c // 2), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_pyr_down_gauss_f32(_p(s), C.c_size_t(0), r, c, _p(d), C.c_size_t(0), _stream()), "ef_op_pyr_down_gauss_f32")
    return _np(d)


def pyrDownUcharGauss(src):
    L = binding.lib()
    s = _dev(src, torch.uint8)
    r, c = s.shape
    d = torch.zeros((r // 2, c // 2), dtype=torch.uint8, device="cuda")
    _chk(L.ef_op_pyr_down_gauss_u8(_p(s), C.c_size_t(0), r, c, _p(d), C.c_size_t(0), _stream()), "ef_op_pyr_down_gauss_u8")
    return _np(d)


def depthBilateral(depth, maxD):
    """ElasticFusion::filterDepth (Shaders/depth_bilateral.frag): raw u16 millimetres -> filtered u16 millimetres."""
    L = binding.lib()
    s = _dev(depth, torch.uint16)
    r, c = s.shape
    d = torch.zeros((r, c), dtype=torch.uint16, device="cuda")
    _chk(L.ef_op_depth_bilateral(_p(s), C.c_size_t(0), r, c, C.c_float(maxD), _p(d), C.c_size_t(0), _stream()), "ef_op_depth_bilateral")
    return _np(d)


def depthMetric(depth, maxD):
    """ElasticFusion::metriciseDepth (Shaders/depth_metric.frag): u16 millimetres -> float metres."""
    L = binding.lib()
    s = _dev(depth, torch.uint16)
    r, c = s.shape
    d = torch.zeros((r, c), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_depth_metric(_p(s), C.c_size_t(0), r, c, C.c_float(maxD), _p(d), C.c_size_t(0), _stream()), "ef_op_depth_metric")
    return _np(d)


def imageBGRToIntensity(rgba):
    L = binding.lib()
    s = _dev(rgba, torch.uint8)
    r, c = s.shape[:2]
    d = torch.zeros((r, c), dtype=torch.uint8, device="cuda")
    _chk(L.ef_op_bgr_to_intensity(_p(s), C.c_size_t(0), r, c, _p(d), C.c_size_t(0), _stream()), "ef_op_bgr_to_intensity")
    return _np(d)


def computeDerivativeImages(img):
    L = binding.lib()
    s = _dev(img, torch.uint8)
    r, c = s.shape
    dx = torch.zeros((r, c), dtype=torch.int16, device="cuda")
    dy = torch.zeros((r, c), dtype=torch.int16, device="cuda")
    _chk(L.ef_op_derivative_images(_p(s), C.c_size_t(0), r, c, _p(dx), _p(dy), C.c_size_t(0), _stream()), "ef_op_derivative_images")
    return _np(dx), _np(dy)


def projectToPointCloud(depth, fx, fy, cx, cy, level=0):
    L = binding.lib()
    d = _dev(depth, torch.float32)
    r, c = d.shape
    cl = torch.zeros((r, c, 3), dtype=torch.float32, device="cuda")
    _chk(L.ef_op_project_point_cloud(_p(d), C.c_size_t(0), r, c, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                     int(level), _p(cl), C.c_size_t(0), _stream()), "ef_op_project_point_cloud")
    return _np(cl)


def icpStep(Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, fx, fy, cx, cy, vmap_g_prev, nmap_g_prev, distThres, angleThres):
    """fx..cy = level intrinsics.  Returns (A 6x6, b 6, residual 2) float32."""
    L = binding.lib()
    vc, nc = _dev(vmap_curr, torch.float32), _dev(nmap_curr, torch.float32)
    vp, npv = _dev(vmap_g_prev, torch.float32), _dev(nmap_g_prev, torch.float32)
    r, c = vc.shape[0] // 3, vc.shape[1]
    sc = _scratch()
    A, b, res = np.zeros(36, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32)
    Rc, tc, Rp, tp = _f9(Rcurr), _f9(tcurr), _f9(Rprev_inv), _f9(tprev)
    vp_ = lambda a: a.ctypes.data_as(C.c_void_p)
    _chk(L.ef_op_icp_step(vp_(Rc), vp_(tc), _p(vc), _p(nc), vp_(Rp), vp_(tp), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                          _p(vp), _p(npv), C.c_size_t(0), C.c_float(distThres), C.c_float(angleThres), r, c, _p(sc), vp_(A), vp_(b), vp_(res),
                          _stream()), "ef_op_icp_step")
    return A.reshape(6, 6), b, res


def computeRgbResidual(minScale, dIdx, dIdy, lastDepth, nextDepth, lastImage, nextImage, maxDepthDelta, kt, krkinv):
    """Returns (corres [rows, cols] DATA_TERM records, sigmaSum, count)."""
    L = binding.lib()
    dx, dy = _dev(dIdx, torch.int16), _dev(dIdy, torch.int16)
    ld, nd = _dev(lastDepth, torch.float32), _dev(nextDepth, torch.float32)
    li, ni = _dev(lastImage, torch.uint8), _dev(nextImage, torch.uint8)
    r, c = dx.shape
    cor = torch.zeros((r, c, 16), dtype=torch.uint8, device="cuda")
    sc = _scratch()
    sig, cnt = C.c_int(0), C.c_int(0)
    ktv, kk = _f9(kt), _f9(krkinv)
    _chk(L.ef_op_rgb_residual(C.c_float(minScale), _p(dx), _p(dy), C.c_size_t(0), _p(ld), _p(nd), C.c_size_t(0), _p(li), _p(ni),
                              C.c_size_t(0), _p(cor), C.c_float(maxDepthDelta), ktv.ctypes.data_as(C.c_void_p),
                              kk.ctypes.data_as(C.c_void_p), r, c, _p(sc), C.byref(sig), C.byref(cnt), _stream()), "ef_op_rgb_residual")
    rec = cor.cpu().numpy().view(DATA_TERM).reshape(r, c)
    return rec, sig.value, cnt.value


def rgbStep(corres, sigma, cloud, fx, fy, dIdx, dIdy, sobelScale):
    """fx, fy = level intrinsics.  Returns (A 6x6, b 6)."""
    L = binding.lib()
    rec = np.ascontiguousarray(corres)
    r, c = rec.shape
    cor = torch.from_numpy(rec.view(np.uint8).reshape(r, c, 16)).to("cuda")
    cl = _dev(cloud, torch.float32)
    dx, dy = _dev(dIdx, torch.int16), _dev(dIdy, torch.int16)
    sc = _scratch()
    A, b = np.zeros(36, np.float32), np.zeros(6, np.float32)
    _chk(L.ef_op_rgb_step(_p(cor), C.c_float(sigma), _p(cl), C.c_size_t(0), C.c_float(fx), C.c_float(fy), _p(dx), _p(dy), C.c_size_t(0),
                          C.c_float(sobelScale), r, c, _p(sc), A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), _stream()),
         "ef_op_rgb_step")
    return A.reshape(6, 6), b


def so3Step(lastImage, nextImage, imageBasis, kinv, krlr):
    L = binding.lib()
    li, ni = _dev(lastImage, torch.uint8), _dev(nextImage, torch.uint8)
    r, c = li.shape
    sc = _scratch()
    A, b, res = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32)
    H, ki, kr = _f9(imageBasis), _f9(kinv), _f9(krlr)
    vp_ = lambda a: a.ctypes.data_as(C.c_void_p)
    _chk(L.ef_op_so3_step(_p(li), _p(ni), C.c_size_t(0), vp_(H), vp_(ki), vp_(kr), r, c, _p(sc), vp_(A), vp_(b), vp_(res), _stream()),
         "ef_op_so3_step")
    return A.reshape(3, 3), b, res


# ---- the step around the tracker: IndexMap::combinedPredict + FillIn, OpenGL in the reference ----
def splatPredict(surfels, pose, cx, cy, fx, fy, rows, cols, maxDepth, confThreshold, time, maxTime, timeDelta):
    """IndexMap::combinedPredict (IndexMap.cpp:468-575): surfels (N, stride/4 >= 12) float32 -- position|confidence,
    colour|instance|initTime|time, normal|radius -- drawn from `pose` (4x4, camera-to-world).  Returns
    (image rgba8, vertex rgba32f, normal rgba32f, time u16) as numpy."""
    L = binding.lib()
    s = _dev(np.ascontiguousarray(surfels, np.float32) if not isinstance(surfels, torch.Tensor) else surfels, torch.float32)
    n, stride = (s.shape[0], s.shape[1] * 4) if s.numel() else (0, 48)
    t_inv = np.ascontiguousarray(np.linalg.inv(np.asarray(pose, np.float64)).astype(np.float32).reshape(16))
    keys = torch.empty(L.ef_op_splat_scratch_bytes(rows, cols), dtype=torch.uint8, device="cuda")
    img = torch.empty((rows, cols, 4), dtype=torch.uint8, device="cuda")
    v = torch.empty((rows, cols, 4), dtype=torch.float32, device="cuda")
    nm = torch.empty((rows, cols, 4), dtype=torch.float32, device="cuda")
    tm = torch.empty((rows, cols), dtype=torch.uint16, device="cuda")
    _chk(L.ef_op_splat_predict(_p(s) if n else None, C.c_size_t(stride), n, t_inv.ctypes.data_as(C.c_void_p), C.c_float(cx), C.c_float(cy),
                               C.c_float(fx), C.c_float(fy), rows, cols, C.c_float(maxDepth), C.c_float(confThreshold), int(time), int(maxTime),
                               int(timeDelta), _p(keys), _p(img), _p(v), _p(nm), _p(tm), _stream()), "ef_op_splat_predict")
    return _np(img), _np(v), _np(nm), _np(tm)


def _fill_geom(name, predicted, depth, cx, cy, fx, fy, passthrough):
    L = binding.lib()
    p = _dev(predicted, torch.float32)
    d = _dev(depth, torch.uint16)
    r, c = d.shape
    out = torch.empty_like(p)
    _chk(getattr(L, name)(_p(p), _p(d), r, c, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy), int(passthrough), _p(out), _stream()), name)
    return _np(out)


def fillVertex(predicted, depth, cx, cy, fx, fy, passthrough=False):
    """FillIn::vertex (Shaders/fill_vertex.frag)"""
    return _fill_geom("ef_op_fill_vertex", predicted, depth, cx, cy, fx, fy, passthrough)


def fillNormal(predicted, depth, cx, cy, fx, fy, passthrough=False):
    """FillIn::normal (Shaders/fill_normal.frag)"""
    return _fill_geom("ef_op_fill_normal", predicted, depth, cx, cy, fx, fy, passthrough)


def fillImage(predicted, rgba, passthrough=False):
    """FillIn::image (Shaders/fill_rgb.frag)"""
    L = binding.lib()
    p, r = _dev(predicted, torch.uint8), _dev(rgba, torch.uint8)
    out = torch.empty_like(p)
    _chk(L.ef_op_fill_rgb(_p(p), _p(r), p.shape[0], p.shape[1], int(passthrough), _p(out), _stream()), "ef_op_fill_rgb")
    return _np(out)
