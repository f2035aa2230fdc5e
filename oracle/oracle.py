"""ctypes bindings for the TEST-ONLY checkers (oracle/ef_oracle.c and oracle/_ref/libef_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this
module.  The product package instancefusion_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libef_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libef_ref.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

DATA_TERM = np.dtype([("zero_x", np.int16), ("zero_y", np.int16), ("one_x", np.int16), ("one_y", np.int16),
                      ("diff", np.float32), ("valid", np.uint8), ("pad", np.uint8, (3,))])
assert DATA_TERM.itemsize == 16


class Stats(C.Structure):
    _fields_ = [("last_icp_error", C.c_float), ("last_icp_count", C.c_float),
                ("last_rgb_error", C.c_float), ("last_rgb_count", C.c_float),
                ("last_so3_error", C.c_float), ("last_so3_count", C.c_float),
                ("last_A", C.c_double * 36), ("last_b", C.c_double * 6),
                ("so3_iterations", C.c_int), ("se3_iterations", C.c_int * 3)]

    def as_dict(self):
        return {
            "last_icp_error": self.last_icp_error, "last_icp_count": self.last_icp_count,
            "last_rgb_error": self.last_rgb_error, "last_rgb_count": self.last_rgb_count,
            "last_so3_error": self.last_so3_error, "last_so3_count": self.last_so3_count,
            "last_A": np.array(self.last_A[:]).reshape(6, 6), "last_b": np.array(self.last_b[:]),
            "so3_iterations": self.so3_iterations, "se3_iterations": list(self.se3_iterations[:]),
        }


def build(force: bool = False) -> None:
    """Compile the C restatement (and, where /root/reference exists, the reference CUDA lib)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "ef_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "libef_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_cpu = None
_ref = None


def cpu():
    global _cpu
    if _cpu is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _cpu = C.CDLL(ORACLE_SO)
        _cpu.efo_tracker_create.restype = C.c_void_p
        _cpu.efo_tracker_buffer.restype = C.c_void_p
    return _cpu


def ref():
    """The reference's own CUDA operators (needs a GPU)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libef_ref.so missing: build it where /root/reference exists (make -C oracle)")
        _ref = C.CDLL(REF_SO)
        _ref.efr_tracker_create.restype = C.c_void_p
        _ref.efr_tracker_download.restype = C.c_size_t
    return _ref


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------------------------------------
# per-operator wrappers.  `impl` is "cpu" (C restatement) or "ref" (reference CUDA kernels).
# ------------------------------------------------------------------------------------------------
def pyr_down_u16(src, impl="cpu"):
    src = _c(src, np.uint16)
    r, c = src.shape
    dst = np.zeros((r // 2, c // 2), np.uint16)
    fn = cpu().efo_pyr_down_u16 if impl == "cpu" else ref().efr_pyr_down_u16
    fn(_fp(src), C.c_int(r), C.c_int(c), _fp(dst))
    return dst


def create_vmap(depth, fx, fy, cx, cy, cutoff, impl="cpu"):
    depth = _c(depth, np.uint16)
    r, c = depth.shape
    v = np.zeros((3 * r, c), np.float32)
    fn = cpu().efo_create_vmap if impl == "cpu" else ref().efr_create_vmap
    fn(_fp(depth), C.c_int(r), C.c_int(c), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
       C.c_float(cutoff), _fp(v))
    return v


def create_nmap(vmap, impl="cpu"):
    vmap = _c(vmap, np.float32)
    r, c = vmap.shape[0] // 3, vmap.shape[1]
    n = np.zeros_like(vmap)
    fn = cpu().efo_create_nmap if impl == "cpu" else ref().efr_create_nmap
    fn(_fp(vmap), C.c_int(r), C.c_int(c), _fp(n))
    return n


def transform_maps(vmap, nmap, R, t, impl="cpu"):
    v = _c(vmap, np.float32).copy()
    n = _c(nmap, np.float32).copy()
    r, c = v.shape[0] // 3, v.shape[1]
    R = _c(R, np.float32)
    t = _c(t, np.float32)
    if impl == "cpu":
        cpu().efo_transform_maps(_fp(v), _fp(n), C.c_int(r), C.c_int(c), _fp(R), _fp(t), _fp(v), _fp(n))
    else:
        ref().efr_transform_maps(_fp(v), _fp(n), C.c_int(r), C.c_int(c), _fp(R), _fp(t))
    return v, n


def copy_maps(v4, n4, impl="cpu"):
    v4 = _c(v4, np.float32)
    n4 = _c(n4, np.float32)
    r, c = v4.shape[:2]
    v = np.zeros((3 * r, c), np.float32)
    n = np.zeros((3 * r, c), np.float32)
    fn = cpu().efo_copy_maps if impl == "cpu" else ref().efr_copy_maps
    fn(_fp(v4), _fp(n4), C.c_int(r), C.c_int(c), _fp(v), _fp(n))
    return v, n


def resize_map(m, normalize, impl="cpu"):
    m = _c(m, np.float32)
    r, c = m.shape[0] // 3, m.shape[1]
    out = np.zeros((3 * (r // 2), c // 2), np.float32)
    fn = cpu().efo_resize_map if impl == "cpu" else ref().efr_resize_map
    fn(_fp(m), C.c_int(r), C.c_int(c), _fp(out), C.c_int(int(normalize)))
    return out


def vertices_to_depth(v4, cutoff, impl="cpu"):
    v4 = _c(v4, np.float32)
    r, c = v4.shape[:2]
    d = np.zeros((r, c), np.float32)
    fn = cpu().efo_vertices_to_depth if impl == "cpu" else ref().efr_vertices_to_depth
    fn(_fp(v4), C.c_int(r), C.c_int(c), C.c_float(cutoff), _fp(d))
    return d


def pyr_down_gauss_f32(src, impl="cpu"):
    src = _c(src, np.float32)
    r, c = src.shape
    dst = np.zeros((r // 2, c // 2), np.float32)
    fn = cpu().efo_pyr_down_gauss_f32 if impl == "cpu" else ref().efr_pyr_down_gauss_f32
    fn(_fp(src), C.c_int(r), C.c_int(c), _fp(dst))
    return dst


def pyr_down_gauss_u8(src, impl="cpu"):
    src = _c(src, np.uint8)
    r, c = src.shape
    dst = np.zeros((r // 2, c // 2), np.uint8)
    fn = cpu().efo_pyr_down_gauss_u8 if impl == "cpu" else ref().efr_pyr_down_gauss_u8
    fn(_fp(src), C.c_int(r), C.c_int(c), _fp(dst))
    return dst


def bgr_to_intensity(rgba, impl="cpu"):
    rgba = _c(rgba, np.uint8)
    r, c = rgba.shape[:2]
    dst = np.zeros((r, c), np.uint8)
    fn = cpu().efo_bgr_to_intensity if impl == "cpu" else ref().efr_bgr_to_intensity
    fn(_fp(rgba), C.c_int(r), C.c_int(c), _fp(dst))
    return dst


def derivative_images(img, impl="cpu"):
    img = _c(img, np.uint8)
    r, c = img.shape
    dx = np.zeros((r, c), np.int16)
    dy = np.zeros((r, c), np.int16)
    fn = cpu().efo_derivative_images if impl == "cpu" else ref().efr_derivative_images
    fn(_fp(img), C.c_int(r), C.c_int(c), _fp(dx), _fp(dy))
    return dx, dy


def depth_bilateral(depth, max_depth_m):
    """Shaders/depth_bilateral.frag (CPU restatement only: the reference runs it as GLSL, there is no reference CUDA)."""
    depth = _c(depth, np.uint16)
    r, c = depth.shape
    dst = np.zeros((r, c), np.uint16)
    cpu().efo_depth_bilateral(_fp(depth), C.c_int(r), C.c_int(c), C.c_float(max_depth_m), _fp(dst))
    return dst


def depth_metric(depth, max_depth_m):
    depth = _c(depth, np.uint16)
    r, c = depth.shape
    dst = np.zeros((r, c), np.float32)
    cpu().efo_depth_metric(_fp(depth), C.c_int(r), C.c_int(c), C.c_float(max_depth_m), _fp(dst))
    return dst


def project_point_cloud(depth, fx, fy, cx, cy, level=0, impl="cpu"):
    """fx..cy are LEVEL-0 intrinsics; `level` divides them by 2^level like CameraModel::operator()."""
    depth = _c(depth, np.float32)
    r, c = depth.shape
    cloud = np.zeros((r, c, 3), np.float32)
    if impl == "cpu":
        d = float(1 << level)
        cpu().efo_project_point_cloud(_fp(depth), C.c_int(r), C.c_int(c), C.c_float(np.float32(fx) / np.float32(d)),
                                      C.c_float(np.float32(fy) / np.float32(d)), C.c_float(np.float32(cx) / np.float32(d)),
                                      C.c_float(np.float32(cy) / np.float32(d)), _fp(cloud))
    else:
        ref().efr_project_point_cloud(_fp(depth), C.c_int(r), C.c_int(c), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                                      C.c_float(cy), C.c_int(level), _fp(cloud))
    return cloud


def unpack_se3(out29):
    out29 = _c(out29, np.float32)
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    res = np.zeros(2, np.float32)
    cpu().efo_unpack_se3(_fp(out29), _fp(A), _fp(b), _fp(res))
    return A.reshape(6, 6), b, res


def unpack_so3(out11):
    out11 = _c(out11, np.float32)
    A = np.zeros(9, np.float32)
    b = np.zeros(3, np.float32)
    res = np.zeros(2, np.float32)
    cpu().efo_unpack_so3(_fp(out11), _fp(A), _fp(b), _fp(res))
    return A.reshape(3, 3), b, res


def icp_step(Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, fx, fy, cx, cy, vmap_g_prev, nmap_g_prev,
             dist_thresh, angle_thresh, impl="cpu", threads=128, blocks=112):
    """returns (A 6x6 f32, b 6 f32, residual[2] f32)"""
    vc, nc = _c(vmap_curr, np.float32), _c(nmap_curr, np.float32)
    vp, npv = _c(vmap_g_prev, np.float32), _c(nmap_g_prev, np.float32)
    r, c = vc.shape[0] // 3, vc.shape[1]
    Rc, tc, Rp, tp = _c(Rcurr, np.float32), _c(tcurr, np.float32), _c(Rprev_inv, np.float32), _c(tprev, np.float32)
    if impl == "cpu":
        out = np.zeros(29, np.float32)
        cpu().efo_icp_step(_fp(Rc), _fp(tc), _fp(vc), _fp(nc), _fp(Rp), _fp(tp), C.c_float(fx), C.c_float(fy),
                           C.c_float(cx), C.c_float(cy), _fp(vp), _fp(npv), C.c_float(dist_thresh),
                           C.c_float(angle_thresh), C.c_int(r), C.c_int(c), _fp(out))
        return unpack_se3(out)
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    res = np.zeros(2, np.float32)
    ref().efr_icp_step(_fp(Rc), _fp(tc), _fp(vc), _fp(nc), _fp(Rp), _fp(tp), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                       C.c_float(cy), _fp(vp), _fp(npv), C.c_float(dist_thresh), C.c_float(angle_thresh), C.c_int(r),
                       C.c_int(c), C.c_int(threads), C.c_int(blocks), _fp(A), _fp(b), _fp(res))
    return A.reshape(6, 6), b, res


def rgb_residual(min_scale, dIdx, dIdy, last_depth, next_depth, last_image, next_image, max_depth_delta, kt, krkinv,
                 impl="cpu", threads=256, blocks=336):
    """returns (corres[rows, cols] DATA_TERM, sigma_sum, count)"""
    dx, dy = _c(dIdx, np.int16), _c(dIdy, np.int16)
    ld, nd = _c(last_depth, np.float32), _c(next_depth, np.float32)
    li, ni = _c(last_image, np.uint8), _c(next_image, np.uint8)
    r, c = dx.shape
    cor = np.zeros((r, c), DATA_TERM)
    sig = C.c_int(0)
    cnt = C.c_int(0)
    ktv, kk = _c(kt, np.float32), _c(krkinv, np.float32)
    if impl == "cpu":
        cpu().efo_rgb_residual(C.c_float(min_scale), _fp(dx), _fp(dy), _fp(ld), _fp(nd), _fp(li), _fp(ni), _fp(cor),
                               C.c_float(max_depth_delta), _fp(ktv), _fp(kk), C.c_int(r), C.c_int(c), C.byref(sig),
                               C.byref(cnt))
    else:
        ref().efr_rgb_residual(C.c_float(min_scale), _fp(dx), _fp(dy), _fp(ld), _fp(nd), _fp(li), _fp(ni), _fp(cor),
                               C.c_float(max_depth_delta), _fp(ktv), _fp(kk), C.c_int(r), C.c_int(c), C.c_int(threads),
                               C.c_int(blocks), C.byref(sig), C.byref(cnt))
    return cor, sig.value, cnt.value


def rgb_step(corres, sigma, cloud, fx, fy, dIdx, dIdy, sobel_scale, impl="cpu", threads=128, blocks=112):
    """fx, fy are the LEVEL intrinsics.  returns (A, b)"""
    cor = np.ascontiguousarray(corres)
    cl = _c(cloud, np.float32)
    dx, dy = _c(dIdx, np.int16), _c(dIdy, np.int16)
    r, c = dx.shape
    if impl == "cpu":
        out = np.zeros(29, np.float32)
        cpu().efo_rgb_step(_fp(cor), C.c_float(sigma), _fp(cl), C.c_float(fx), C.c_float(fy), _fp(dx), _fp(dy),
                           C.c_float(sobel_scale), C.c_int(r), C.c_int(c), _fp(out))
        A, b, _ = unpack_se3(out)
        return A, b
    A = np.zeros(36, np.float32)
    b = np.zeros(6, np.float32)
    ref().efr_rgb_step(_fp(cor), C.c_float(sigma), _fp(cl), C.c_float(fx), C.c_float(fy), _fp(dx), _fp(dy),
                       C.c_float(sobel_scale), C.c_int(r), C.c_int(c), C.c_int(threads), C.c_int(blocks), _fp(A), _fp(b))
    return A.reshape(6, 6), b


def so3_step(last_image, next_image, image_basis, kinv, krlr, impl="cpu", threads=160, blocks=64):
    li, ni = _c(last_image, np.uint8), _c(next_image, np.uint8)
    r, c = li.shape
    H, ki, kr = _c(image_basis, np.float32), _c(kinv, np.float32), _c(krlr, np.float32)
    if impl == "cpu":
        out = np.zeros(11, np.float32)
        cpu().efo_so3_step(_fp(li), _fp(ni), _fp(H), _fp(ki), _fp(kr), C.c_int(r), C.c_int(c), _fp(out))
        return unpack_so3(out)
    A = np.zeros(9, np.float32)
    b = np.zeros(3, np.float32)
    res = np.zeros(2, np.float32)
    ref().efr_so3_step(_fp(li), _fp(ni), _fp(H), _fp(ki), _fp(kr), C.c_int(r), C.c_int(c), C.c_int(threads),
                       C.c_int(blocks), _fp(A), _fp(b), _fp(res))
    return A.reshape(3, 3), b, res


# ------------------------------------------------------------------------------------------------
# tracker wrapper: same call sequence as RGBDOdometry (RGBDOdometry.h:35-73)
# ------------------------------------------------------------------------------------------------
_BUF_DTYPES = {"vmap_curr": (np.float32, 3), "nmap_curr": (np.float32, 3), "vmap_g_prev": (np.float32, 3),
               "nmap_g_prev": (np.float32, 3), "last_depth": (np.float32, 1), "next_depth": (np.float32, 1),
               "last_image": (np.uint8, 1), "next_image": (np.uint8, 1), "last_next_image": (np.uint8, 1),
               "dIdx": (np.int16, 1), "dIdy": (np.int16, 1), "depth_tmp": (np.uint16, 1)}


class OracleTracker:
    """impl="cpu": C restatement.  impl="ref": reference CUDA operators under the restated host loop."""

    def __init__(self, width, height, cx, cy, fx, fy, dist_thresh=0.10,
                 angle_thresh=float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0))), impl="cpu"):
        self.impl = impl
        self.w, self.h = width, height
        self.lib = cpu() if impl == "cpu" else ref()
        self.p = "efo_" if impl == "cpu" else "efr_"
        self.t = C.c_void_p(getattr(self.lib, self.p + "tracker_create")(
            C.c_int(width), C.c_int(height), C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy),
            C.c_float(dist_thresh), C.c_float(angle_thresh)))
        self.stats = Stats()

    def close(self):
        if self.t:
            getattr(self.lib, self.p + "tracker_destroy")(self.t)
            self.t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args, device=False):
        fn = getattr(self.lib, self.p + name)
        if self.impl == "ref":
            fn(self.t, *args, C.c_int(int(device)))
        else:
            fn(self.t, *args)

    def init_icp_depth(self, depth, cutoff):
        self._call("init_icp_depth", _fp(_c(depth, np.uint16)), C.c_float(cutoff))

    def init_icp_maps(self, v4, n4, cutoff):
        self._call("init_icp_maps", _fp(_c(v4, np.float32)), _fp(_c(n4, np.float32)), C.c_float(cutoff))

    def init_icp_model(self, v4, n4, cutoff, pose):
        self._call("init_icp_model", _fp(_c(v4, np.float32)), _fp(_c(n4, np.float32)), C.c_float(cutoff),
                   _fp(_c(pose, np.float32)))

    def init_rgb(self, rgba):
        self._call("init_rgb", _fp(_c(rgba, np.uint8)))

    def init_rgb_model(self, rgba):
        self._call("init_rgb_model", _fp(_c(rgba, np.uint8)))

    def init_first_rgb(self, rgba):
        self._call("init_first_rgb", _fp(_c(rgba, np.uint8)))

    # raw device-pointer variants (ref impl only; used by bench.py --impl reference)
    def call_dev(self, name, *args):
        getattr(self.lib, self.p + name)(self.t, *args, C.c_int(1))

    def get_incremental_transformation(self, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3):
        tr = _c(trans, np.float32).copy()
        ro = _c(rot, np.float32).reshape(9).copy()
        getattr(self.lib, self.p + "get_incremental_transformation")(
            self.t, _fp(tr), _fp(ro), C.c_int(int(rgb_only)), C.c_float(icp_weight), C.c_int(int(pyramid)),
            C.c_int(int(fast_odom)), C.c_int(int(so3)), C.byref(self.stats))
        return tr, ro.reshape(3, 3), self.stats.as_dict()

    def buffer(self, name, level):
        dt, planes = _BUF_DTYPES[name]
        r, c = (self.h >> level) * planes, self.w >> level
        if self.impl == "cpu":
            ptr = self.lib.efo_tracker_buffer(self.t, name.encode(), C.c_int(level))
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(r, c))
            return arr.copy()
        out = np.zeros((r, c), dt)
        n = self.lib.efr_tracker_download(self.t, name.encode(), C.c_int(level), _fp(out))
        assert n == out.nbytes, (name, n, out.nbytes)
        return out


# ---- model prediction + fill-in (OpenGL in the reference; CPU restatement only, parity against GL unpinned) ----
def splat_predict(surfels, pose, cx, cy, fx, fy, rows, cols, max_depth, conf_threshold, time, max_time, time_delta):
    s = _c(surfels, np.float32)
    n, stride = (s.shape[0], s.shape[1]) if s.size else (0, 12)
    t_inv = np.ascontiguousarray(np.linalg.inv(np.asarray(pose, np.float64)).astype(np.float32).reshape(16))
    img = np.zeros((rows, cols, 4), np.uint8)
    v = np.zeros((rows, cols, 4), np.float32)
    nm = np.zeros((rows, cols, 4), np.float32)
    tm = np.zeros((rows, cols), np.uint16)
    cpu().efo_splat_predict(_fp(s) if n else None, C.c_int(stride), C.c_int(n), _fp(t_inv), C.c_float(cx), C.c_float(cy), C.c_float(fx),
                            C.c_float(fy), C.c_int(rows), C.c_int(cols), C.c_float(max_depth), C.c_float(conf_threshold), C.c_int(time),
                            C.c_int(max_time), C.c_int(time_delta), _fp(img), _fp(v), _fp(nm), _fp(tm))
    return img, v, nm, tm


def fill_vertex(predicted, depth, cx, cy, fx, fy, passthrough=False, normal=False):
    p, d = _c(predicted, np.float32), _c(depth, np.uint16)
    r, c = d.shape
    out = np.zeros_like(p)
    fn = cpu().efo_fill_normal if normal else cpu().efo_fill_vertex
    fn(_fp(p), _fp(d), C.c_int(r), C.c_int(c), C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy), C.c_int(int(passthrough)), _fp(out))
    return out


def fill_normal(predicted, depth, cx, cy, fx, fy, passthrough=False):
    return fill_vertex(predicted, depth, cx, cy, fx, fy, passthrough, normal=True)


def fill_rgb(predicted, raw, passthrough=False):
    p, r = _c(predicted, np.uint8), _c(raw, np.uint8)
    out = np.zeros_like(p)
    cpu().efo_fill_rgb(_fp(p), _fp(r), C.c_int(p.shape[0]), C.c_int(p.shape[1]), C.c_int(int(passthrough)), _fp(out))
    return out
