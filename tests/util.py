"""Shared test inputs: seeded synthetic frames (instancefusion_b200.synth) and adversarial images."""
from __future__ import annotations

import ctypes
import functools
import math

import numpy as np

from instancefusion_b200 import synth


def u16(t):
    import torch
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


@functools.lru_cache(maxsize=8)
def frame_pair(width=640, height=480, seed=1234):
    """(K, pose0, pose1, f0, f1) with numpy arrays; config 1 of BASELINE.json."""
    K = synth.Intrinsics.kinect(width, height)
    pose0, pose1, f0, f1 = synth.frame_pair(K, seed=seed)

    def to_np(f):
        return {"depth": u16(f["depth"]), "rgba": f["rgba"].numpy(), "vmap": f["vmap"].numpy(), "nmap": f["nmap"].numpy()}

    return K, pose0.numpy(), pose1.numpy(), to_np(f0), to_np(f1)


def punch_holes(depth, seed=7, frac=0.08):
    """zero out random blobs + a border stripe: exercises every invalid-pixel branch."""
    rng = np.random.default_rng(seed)
    d = depth.copy()
    h, w = d.shape
    n = int(frac * h * w / 64)
    ys = rng.integers(0, h - 8, n)
    xs = rng.integers(0, w - 8, n)
    for y, x in zip(ys, xs):
        d[y:y + 8, x:x + 8] = 0
    d[:, :3] = 0
    d[h - 2:, :] = 0
    return d


def holes_in_maps(vmap, nmap, seed=11):
    rng = np.random.default_rng(seed)
    v, n = vmap.copy(), nmap.copy()
    h, w = v.shape[:2]
    for _ in range(40):
        y, x = rng.integers(0, h - 6), rng.integers(0, w - 6)
        v[y:y + 6, x:x + 6] = 0
        n[y:y + 6, x:x + 6] = 0
    return v, n


def random_image_u8(h, w, seed=3, zero_frac=0.05):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img[rng.random((h, w)) < zero_frac] = 0
    return img


def ulp_diff(a, b):
    """max ULP distance between two float32 arrays (NaN == NaN)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2147483648) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2147483648) - bi, bi)
    d = np.abs(ai - bi)
    d[both_nan] = 0
    return d


def masked_map_compare(got, want, rows):
    """3-plane maps: x plane must agree incl. NaN pattern; y,z compared only where x is valid
    (invalid pixels keep stale y/z in the reference).  Returns max ulp."""
    gx, wx = got[:rows], want[:rows]
    assert np.array_equal(np.isnan(gx), np.isnan(wx)), "NaN pattern of the x plane differs"
    valid = ~np.isnan(wx)
    m = 0
    for c in range(3):
        g, w = got[c * rows:(c + 1) * rows][valid], want[c * rows:(c + 1) * rows][valid]
        if g.size:
            m = max(m, int(ulp_diff(g, w).max()))
    return m


def rot_err(Ra, Rb):
    """rotation angle (rad) of Ra * Rb^T, computed from the skew part so it stays accurate near 0
    (acos(1 - eps) would amplify float32 rounding of the inputs to ~3e-4)."""
    R = np.asarray(Ra, np.float64) @ np.asarray(Rb, np.float64).T
    s = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = (np.trace(R) - 1) / 2
    return math.atan2(float(np.linalg.norm(s)), float(c))


def se3_level_params(K, level):
    d = np.float32(1 << level)
    return (np.float32(K.fx) / d, np.float32(K.fy) / d, np.float32(K.cx) / d, np.float32(K.cy) / d)


# ------------------------------------------------------------------------------------------------
# the reference's real-data fixture (elasticfusionpublic/GPUTest/{1c,1d,2c,2d}.png, decoded by
# tests/golden/make_gputest_fixture.py) turned into tracker inputs the way GPUTest.cpp does
# ------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=1)
def gputest_inputs():
    """(K, f0, f1): f0 = model maps + colour from frame 1 (loadVertices / loadImage, GPUTest.cpp:62-130, :30-40),
    f1 = depth in millimetres + colour of frame 2 (loadDepth: raw / 5, :42-60).  K = 528, 528, 320, 240 (:150-152)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gputest_pair.npz"))
    K = synth.Intrinsics(640, 480, 528.0, 528.0, 320.0, 240.0)
    d1 = z["depth1_raw"]
    f32 = np.float32
    depth = d1.astype(f32) / f32(5000.0)
    rows, cols = np.meshgrid(np.arange(480, dtype=f32), np.arange(640, dtype=f32), indexing="ij")
    inv_fx, inv_fy = f32(1.0) / f32(K.fx), f32(1.0) / f32(K.fy)
    vx = (cols - f32(K.cx)) * depth * inv_fx  # getVertex, :62-67 (float arithmetic)
    vy = (rows - f32(K.cy)) * depth * inv_fy
    V = np.stack([vx, vy, depth], axis=-1).astype(f32)
    ok = np.zeros((480, 640), bool)
    c = d1 > 0
    ok[1:-1, 1:-1] = c[1:-1, 1:-1] & c[2:, 1:-1] & c[1:-1, 2:] & c[:-2, 1:-1] & c[1:-1, :-2]  # :86-90
    del_x = np.zeros_like(V)
    del_y = np.zeros_like(V)
    del_x[:, :-1] = V[:, 1:] - V[:, :-1]
    del_y[:-1, :] = V[1:, :] - V[:-1, :]
    n = np.cross(del_x, del_y).astype(f32)
    nn = np.sqrt((n * n).sum(-1, dtype=f32)).astype(f32)
    n = (n / np.where(nn > 0, nn, f32(1))[..., None]).astype(f32)
    vmap = np.zeros((480, 640, 4), f32)
    nmap = np.zeros((480, 640, 4), f32)
    vmap[..., 3] = 1
    nmap[..., 3] = 1
    vmap[..., :3] = np.where(ok[..., None], V, 0)
    nmap[..., :3] = np.where(ok[..., None], n, 0)
    # (the border row / column of GPUTest's Img buffers is never written; zeros = "no data" here)
    vmap[0], vmap[-1], vmap[:, 0], vmap[:, -1] = 0, 0, 0, 0
    nmap[0], nmap[-1], nmap[:, 0], nmap[:, -1] = 0, 0, 0, 0

    def rgba(rgb):
        out = np.full((480, 640, 4), 255, np.uint8)  # GL_RGB upload into an RGBA texture: alpha = 1
        out[..., :3] = rgb
        return out

    f0 = {"vmap": vmap, "nmap": nmap, "rgba": rgba(z["rgb1"]), "depth": (d1 // 5).astype(np.uint16)}
    f1 = {"depth": (z["depth2_raw"] // 5).astype(np.uint16), "rgba": rgba(z["rgb2"])}
    return K, f0, f1


# ------------------------------------------------------------------------------------------------
# cudaArray inputs (what a GL-interop caller hands to the *_array entry points): CUDA runtime through ctypes
# ------------------------------------------------------------------------------------------------
class _ChannelDesc(ctypes.Structure):  # cudaChannelFormatDesc
    _fields_ = [("x", ctypes.c_int), ("y", ctypes.c_int), ("z", ctypes.c_int), ("w", ctypes.c_int), ("f", ctypes.c_int)]


@functools.lru_cache(maxsize=1)
def _cudart():
    import ctypes as C
    import glob
    import os
    import torch
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    cands += ["/usr/local/cuda/lib64/libcudart.so", "libcudart.so.12", "libcudart.so"]
    for c in cands:
        try:
            return C.CDLL(c)
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


class CudaArray:
    """a 2-D cudaArray holding a host image: kind in {"u16", "rgba8", "rgba32f"} (the three texture formats the
    tracker is fed with, GPUTexture / RGBDOdometry.cpp:120-128)."""
    FORMATS = {"u16": ((16, 0, 0, 0, 1), np.uint16, 1), "rgba8": ((8, 8, 8, 8, 1), np.uint8, 4), "rgba32f": ((32, 32, 32, 32, 2), np.float32, 4)}

    def __init__(self, host, kind):
        import ctypes as C
        bits, dt, ch = self.FORMATS[kind]
        a = np.ascontiguousarray(host, dt)
        h, w = a.shape[0], a.shape[1]
        assert a.size == h * w * ch
        rt = _cudart()
        desc = _ChannelDesc(*bits)
        self.arr = C.c_void_p()
        rc = rt.cudaMallocArray(C.byref(self.arr), C.byref(desc), C.c_size_t(w), C.c_size_t(h), C.c_uint(0))
        assert rc == 0, ("cudaMallocArray", rc)
        row = w * ch * a.itemsize
        rc = rt.cudaMemcpy2DToArray(self.arr, C.c_size_t(0), C.c_size_t(0), a.ctypes.data_as(C.c_void_p), C.c_size_t(row), C.c_size_t(row),
                                    C.c_size_t(h), C.c_int(1))  # cudaMemcpyHostToDevice
        assert rc == 0, ("cudaMemcpy2DToArray", rc)
        rc = rt.cudaDeviceSynchronize()
        assert rc == 0

    @property
    def cuda_array_handle(self):
        """the cudaArray_t as an integer: what instancefusion_b200.RGBDOdometry routes to the *_array entry points"""
        return self.arr.value

    def free(self):
        if self.arr:
            _cudart().cudaFreeArray(self.arr)
            self.arr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# the reference tracker together with its own launch-shape spread
# ------------------------------------------------------------------------------------------------
class RefEnsemble:
    """The reference's CUDA tracker at GPUConfig's default launch shapes (the parity target) plus the same tracker at other
    (threads, blocks) shapes.  The reference's reductions add floats in an order that depends on the launch shape
    (reduce.cu:90-255), and a coarse-to-fine Gauss-Newton with fixed iteration counts amplifies that on some
    frames: measured on B200 (tools/parity_spread.py, profiles/r02_parity_spread.txt) the reference moves by up to 4e-4 m
    against ITSELF on single frames of BASELINE's trajectories while most frames stay below 2e-6.  A parity test can
    therefore ask for 1e-5 -- or, on the frames where the reference itself is not reproducible to 1e-5, for its own
    spread."""
    ALT = ((256, 96), (96, 148), (512, 32), (128, 296), (64, 64), (384, 56))
    K = 3.0  # bound = K x the spread seen over these shapes (a small sample: a lower bound of what the reference can do to itself)

    def __init__(self, w, h, K):
        from oracle import oracle as O
        self.refs = [O.OracleTracker(w, h, K.cx, K.cy, K.fx, K.fy, impl="ref") for _ in range(1 + len(self.ALT))]
        for r, (t, b) in zip(self.refs[1:], self.ALT):
            r.lib.efr_tracker_set_config(r.t, t, b, t, b, t, b, t, b)
        self.ref = self.refs[0]

    def each(self, fn):
        for r in self.refs:
            fn(r)

    def track(self, trans, rot, **kw):
        """-> (t, R, stats) of the default shape and the spread {"t", "r", "A", "b"} over the other shapes"""
        outs = [r.get_incremental_transformation(trans, rot, **kw) for r in self.refs]
        t0, R0, st0 = outs[0]
        nA = float(np.linalg.norm(st0["last_A"]))
        spread = {"t": 0.0, "r": 0.0, "A": 0.0, "b": 0.0, "icp": 0.0, "rgb": 0.0, "icp_err": 0.0}
        for t, R, st in outs[1:]:
            spread["t"] = max(spread["t"], float(np.abs(t - t0).max()))
            spread["r"] = max(spread["r"], rot_err(R, R0))
            spread["A"] = max(spread["A"], float(np.linalg.norm(st["last_A"] - st0["last_A"])) / max(nA, 1e-30))
            spread["b"] = max(spread["b"], float(np.linalg.norm(st["last_b"] - st0["last_b"])))
            spread["icp"] = max(spread["icp"], abs(st["last_icp_count"] - st0["last_icp_count"]))
            spread["rgb"] = max(spread["rgb"], abs(st["last_rgb_count"] - st0["last_rgb_count"]))
            if st0["last_icp_error"] > 0:
                spread["icp_err"] = max(spread["icp_err"], abs(st["last_icp_error"] - st0["last_icp_error"]) / st0["last_icp_error"])
        return t0, R0, st0, spread

    def close(self):
        for r in self.refs:
            r.close()
