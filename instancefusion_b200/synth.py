"""Deterministic synthetic RGB-D data for the tracker.

A ray-cast axis-aligned box room (6 x 3 x 5 m) with three interior boxes, seen by a pinhole
camera.  For every pose the generator emits exactly the tracker's input formats:

  depth   uint16 [H, W]      millimetres, 0 = invalid (beyond 8 m)      -> initICP(depth)
  rgba    uint8  [H, W, 4]   procedural texture, alpha 255              -> initRGB / initRGBModel
  vmap    float32 [H, W, 4]  camera-frame vertex (x, y, z, 1), zeros if invalid  -> initICPModel
  nmap    float32 [H, W, 4]  camera-frame normal (nx, ny, nz, 1), same convention as
                             computeNmapKernel (cudafuncs.cu:180: cross(right, down), i.e. pointing
                             away from the camera)

The reference reads these from GL textures (GPUTexture); here they are plain tensors.  Everything is
written with torch ops so the same code runs on CPU (tests) and on the GPU (bench.py fills HBM-resident
sequences without a host round trip).  Noise is a counter-based integer hash, so a frame is a pure
function of (seed, frame index, pixel).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

ROOM = (6.0, 3.0, 5.0)
# interior boxes: (min xyz, max xyz)
BOXES = (
    ((1.0, 1.8, 1.2), (2.2, 3.0, 2.0)),
    ((3.6, 2.1, 2.6), (4.6, 3.0, 3.8)),
    ((2.4, 2.4, 3.9), (3.3, 3.0, 4.6)),
)
MAX_RANGE = 8.0
# residual depth noise after the (out-of-scope) bilateral filter: sigma = NOISE_MM * z^2 millimetres
NOISE_MM = 0.1


@dataclass
class Intrinsics:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    @staticmethod
    def kinect(width: int = 640, height: int = 480) -> "Intrinsics":
        # K = (528, 528, 320, 240) at 640x480 (GPUTest.cpp:150-152), same field of view at other sizes
        s = width / 640.0
        return Intrinsics(width, height, 528.0 * s, 528.0 * s, width / 2.0, height / 2.0)


def _hash_u32(x: torch.Tensor) -> torch.Tensor:
    """lowbias32-style integer hash on int64 tensors holding uint32 values."""
    m = 0xFFFFFFFF
    x = x & m
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & m
    x = ((x ^ (x >> 15)) * 0x846CA68B) & m
    x = x ^ (x >> 16)
    return x & m


def _hash_unit(ix: torch.Tensor, iy: torch.Tensor, iz: torch.Tensor, seed: int) -> torch.Tensor:
    h = _hash_u32(ix * 73856093 + iy * 19349663 + iz * 83492791 + seed * 2654435761)
    return h.to(torch.float64) / 4294967296.0


def exp_se3(xi) -> torch.Tensor:
    """4x4 (float64) from a twist (tx, ty, tz, rx, ry, rz): rotation = Rodrigues(r), translation = t."""
    t = torch.tensor(xi[:3], dtype=torch.float64)
    r = torch.tensor(xi[3:], dtype=torch.float64)
    th = float(torch.linalg.norm(r))
    T = torch.eye(4, dtype=torch.float64)
    if th > 1e-12:
        k = r / th
        K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=torch.float64)
        T[:3, :3] = torch.eye(3, dtype=torch.float64) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
    T[:3, 3] = t
    return T


def look_at(eye, target, roll: float = 0.0) -> torch.Tensor:
    """camera-to-world 4x4 (float64); camera x right, y down, z forward; world y is down."""
    eye = torch.tensor(eye, dtype=torch.float64)
    target = torch.tensor(target, dtype=torch.float64)
    z = target - eye
    z = z / torch.linalg.norm(z)
    down = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    x = torch.linalg.cross(down, z)
    x = x / torch.linalg.norm(x)
    y = torch.linalg.cross(z, x)
    c, s = math.cos(roll), math.sin(roll)
    x2 = c * x + s * y
    y2 = -s * x + c * y
    T = torch.eye(4, dtype=torch.float64)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x2, y2, z, eye
    return T


def trajectory(n_frames: int, seed: int = 2024, speed: float = 1.0, fps: float = 30.0) -> torch.Tensor:
    """Handheld Lissajous trajectory (config 2 of BASELINE.json): [n, 4, 4] float64 camera-to-world poses.

    amplitudes 0.4/0.15/0.3 m, periods 7.3/5.1/9.7 s, look-at with 3 degree roll; `speed` < 1 gives the
    slower dyson_lab-like motion of config 4."""
    ph = [(_hash_u32(torch.tensor(seed * 7919 + k)).item() / 4294967296.0) * 2 * math.pi for k in range(4)]
    poses = []
    for i in range(n_frames):
        t = i / fps * speed
        eye = (
            3.0 + 0.4 * math.sin(2 * math.pi * t / 7.3 + ph[0]),
            1.4 + 0.15 * math.sin(2 * math.pi * t / 5.1 + ph[1]),
            0.9 + 0.3 * math.sin(2 * math.pi * t / 9.7 + ph[2]),
        )
        target = (
            3.0 + 0.5 * math.sin(2 * math.pi * t / 11.0 + ph[3]),
            1.9 + 0.2 * math.sin(2 * math.pi * t / 6.3 + ph[0]),
            4.2,
        )
        roll = math.radians(3.0) * math.sin(2 * math.pi * t / 4.1 + ph[1])
        poses.append(look_at(eye, target, roll))
    return torch.stack(poses)


def _texture(p: torch.Tensor, face: torch.Tensor, seed: int) -> torch.Tensor:
    """Procedural intensity-rich colour at world points p [N,3] on surfaces with id face [N] -> uint8 [N,3]."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    f = face.to(torch.float64)
    # three sinusoids with 6-25 cm wavelengths: enough gradient to pass minimumGradientMagnitudes
    s = (
        0.30 * torch.sin(2 * math.pi * (x + 0.37 * y) / 0.21 + 1.3 * f)
        + 0.25 * torch.sin(2 * math.pi * (z + 0.23 * x) / 0.13 + 0.7 * f)
        + 0.20 * torch.sin(2 * math.pi * (y + 0.41 * z) / 0.077 + 2.1 * f)
    )
    # hashed value noise on a 4 cm lattice (piecewise constant -> sharp edges)
    cell = 0.04
    ix = torch.floor(x / cell).to(torch.int64)
    iy = torch.floor(y / cell).to(torch.int64)
    iz = torch.floor(z / cell).to(torch.int64)
    n = _hash_unit(ix + 1000, iy + 1000, iz + 1000 + 17 * face.to(torch.int64), seed) - 0.5
    base = 0.5 + 0.5 * s + 0.35 * n
    r = base + 0.10 * torch.sin(2 * math.pi * x / 0.9 + f)
    g = base + 0.10 * torch.sin(2 * math.pi * y / 0.7 + 2 * f)
    b = base + 0.10 * torch.sin(2 * math.pi * z / 1.1 + 3 * f)
    rgb = torch.stack([r, g, b], dim=1).clamp(0.02, 1.0)  # never exactly 0: 0 means "no data" to the tracker
    return (rgb * 255.0).round().to(torch.uint8)


def render(pose_c2w: torch.Tensor, K: Intrinsics, seed: int = 1234, frame_id: int = 0, noise: bool = True,
           device: str | torch.device = "cpu") -> dict:
    """Ray-cast one frame.  Returns dict(depth, rgba, vmap, nmap, z) on `device`."""
    dev = torch.device(device)
    H, W = K.height, K.width
    T = pose_c2w.to(dev, torch.float64)
    R, c = T[:3, :3], T[:3, 3]
    v, u = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float64),
                          torch.arange(W, device=dev, dtype=torch.float64), indexing="ij")
    dc = torch.stack([(u - K.cx) / K.fx, (v - K.cy) / K.fy, torch.ones_like(u)], dim=-1).reshape(-1, 3)
    d = dc @ R.T  # world direction; parameter t along it IS the camera-frame depth z
    N = d.shape[0]
    inf = torch.full((N,), float("inf"), device=dev, dtype=torch.float64)
    best_t = inf.clone()
    best_face = torch.zeros(N, device=dev, dtype=torch.int64)
    best_n = torch.zeros(N, 3, device=dev, dtype=torch.float64)
    eps = 1e-12
    dsafe = torch.where(d.abs() < eps, torch.full_like(d, eps), d)

    # room: we are inside, the hit is the nearest exit plane
    lo = torch.zeros(3, device=dev, dtype=torch.float64)
    hi = torch.tensor(ROOM, device=dev, dtype=torch.float64)
    t_lo = (lo - c) / dsafe
    t_hi = (hi - c) / dsafe
    t_exit = torch.where(d > 0, t_hi, t_lo)  # per axis
    t_room, ax = t_exit.min(dim=1)
    sign = torch.where(torch.gather(d, 1, ax[:, None])[:, 0] > 0, 1.0, -1.0)
    best_t = t_room
    best_face = ax * 2 + (sign > 0).to(torch.int64)
    best_n = torch.nn.functional.one_hot(ax, 3).to(torch.float64) * sign[:, None]

    for bi, (bmin, bmax) in enumerate(BOXES):
        bmin_t = torch.tensor(bmin, device=dev, dtype=torch.float64)
        bmax_t = torch.tensor(bmax, device=dev, dtype=torch.float64)
        t0 = (bmin_t - c) / dsafe
        t1 = (bmax_t - c) / dsafe
        tn = torch.minimum(t0, t1)
        tf = torch.maximum(t0, t1)
        t_enter, ax_e = tn.max(dim=1)
        t_leave = tf.min(dim=1).values
        hit = (t_enter < t_leave) & (t_enter > 1e-6) & (t_enter < best_t)
        sgn = torch.where(torch.gather(d, 1, ax_e[:, None])[:, 0] > 0, 1.0, -1.0)  # normal along the ray
        nb = torch.nn.functional.one_hot(ax_e, 3).to(torch.float64) * sgn[:, None]
        best_t = torch.where(hit, t_enter, best_t)
        best_face = torch.where(hit, 6 + bi * 6 + ax_e * 2 + (sgn > 0).to(torch.int64), best_face)
        best_n = torch.where(hit[:, None], nb, best_n)

    z = best_t
    valid = torch.isfinite(z) & (z > 0.05) & (z < MAX_RANGE)
    p_world = c[None, :] + d * z[:, None]
    rgb = _texture(p_world, best_face, seed)
    rgb = torch.where(valid[:, None], rgb, torch.zeros_like(rgb))

    n_cam = best_n @ R  # R^T n, oriented along the viewing ray (away from the camera)
    zf = z.to(torch.float32)
    vm = torch.stack([dc[:, 0].to(torch.float32) * zf, dc[:, 1].to(torch.float32) * zf, zf,
                      torch.ones_like(zf)], dim=1)
    nm = torch.cat([n_cam.to(torch.float32), torch.ones(N, 1, device=dev)], dim=1)
    vm = torch.where(valid[:, None], vm, torch.zeros_like(vm))
    nm = torch.where(valid[:, None], nm, torch.zeros_like(nm))

    zmm = z * 1000.0
    if noise:
        pix = torch.arange(N, device=dev, dtype=torch.int64)
        # sum of 4 uniforms - 2 : zero mean, sigma = sqrt(1/3); scale to sigma = NOISE_MM * z^2
        g = sum(_hash_unit(pix, torch.full_like(pix, frame_id), torch.full_like(pix, k), seed) for k in range(4)) - 2.0
        zmm = zmm + g * math.sqrt(3.0) * NOISE_MM * (z * z)
    depth = torch.where(valid, zmm.round().clamp(0, 65535), torch.zeros_like(zmm)).to(torch.int32).to(torch.uint16)
    rgba = torch.cat([rgb, torch.full((N, 1), 255, device=dev, dtype=torch.uint8)], dim=1)
    return {
        "depth": depth.reshape(H, W),
        "rgba": rgba.reshape(H, W, 4).contiguous(),
        "vmap": vm.reshape(H, W, 4).contiguous(),
        "nmap": nm.reshape(H, W, 4).contiguous(),
    }


def frame_pair(K: Intrinsics, seed: int = 1234, device="cpu"):
    """Config 1: pose0 and pose1 = pose0 * exp(xi), xi = (12, -7, 9 mm; 0.010, -0.006, 0.008 rad)."""
    pose0 = look_at((3.0, 1.4, 0.9), (3.1, 1.9, 4.2), 0.0)
    xi = (0.012, -0.007, 0.009, 0.010, -0.006, 0.008)
    pose1 = pose0 @ exp_se3(xi)
    f0 = render(pose0, K, seed=seed, frame_id=0, device=device)
    f1 = render(pose1, K, seed=seed, frame_id=1, device=device)
    return pose0, pose1, f0, f1


def surfels_from_frame(pose_c2w, vmap, nmap, rgba, K: Intrinsics, time: int = 1, confidence: float = 10.0, stride_floats: int = 12):
    """A surfel map made of ONE frame, the way GlobalModel::initialise seeds the map from the first frame
    (elasticfusionpublic/Core/src/Shaders/init_unstable.vert, surfels.glsl:19-34): one surfel per valid pixel with
    position | confidence, encoded colour | 0 | initTime | time, normal | radius, in the WORLD frame.
    numpy in, numpy (N, stride_floats) float32 out (stride 12 = ElasticFusion's 48-byte vertex, 64 = InstanceFusion's 256).
    Input for the model-prediction operator (ops.splatPredict) in tests and tools; not a fusion step."""
    import numpy as np
    pose = np.asarray(pose_c2w, np.float32)
    v = np.asarray(vmap, np.float32).reshape(-1, 4)[:, :3]
    n = np.asarray(nmap, np.float32).reshape(-1, 4)[:, :3]
    c = np.asarray(rgba, np.uint8).reshape(-1, 4).astype(np.int32)
    ok = (v[:, 2] > 0) & np.isfinite(n[:, 0]) & (np.abs(n).sum(1) > 0)
    v, n, c = v[ok], n[ok], c[ok]
    mean_focal = 0.5 * (K.fx + K.fy)
    radius = (v[:, 2] / mean_focal) * 1.41421356237
    radius = np.minimum(2.0 * radius, radius / np.maximum(np.abs(n[:, 2]), 1e-6))  # surfels.glsl:25-33
    out = np.zeros((v.shape[0], stride_floats), np.float32)
    out[:, 0:3] = v @ pose[:3, :3].T + pose[:3, 3]
    out[:, 3] = confidence
    out[:, 4] = ((c[:, 0] << 16) + (c[:, 1] << 8) + c[:, 2]).astype(np.float32)  # color.glsl:19-25 encodeColor
    out[:, 6] = time
    out[:, 7] = time
    out[:, 8:11] = n @ pose[:3, :3].T
    out[:, 11] = radius
    return out
