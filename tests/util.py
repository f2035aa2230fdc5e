"""Shared test inputs: seeded synthetic frames (instancefusion_b200.synth) and adversarial images."""
from __future__ import annotations

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
