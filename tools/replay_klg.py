#!/usr/bin/env python
"""Replay a .klg RGB-D log through the tracker, GL-free:  klg -> raw depth -> CUDA bilateral filter
(ef_init_icp_depth_raw) -> frame-to-frame joint ICP+RGB tracking.

The reference tracks frame-to-MODEL against the surfel map's prediction (OpenGL, out of scope); without the map the
"model" of frame k here is frame k-1 itself: its vertex / normal maps (back-projected filtered depth, normals by forward
differences like createNMap) at the pose estimated for k-1 -- i.e. closed-loop frame-to-frame odometry, every pose
depends on the previous one.  With --synthetic N a log of the bench's trajectory is written first (Logger2 layout: zlib
depth + JPEG colour) and the estimated trajectory is scored against its ground truth.

    python tools/replay_klg.py --synthetic 120            # writes /tmp/synth.klg, replays it, prints ATE and frames/s
    python tools/replay_klg.py --klg dyson_lab.klg        # any 640x480 ElasticFusion log
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instancefusion_b200 as ef  # noqa: E402
from instancefusion_b200 import rgbd_odometry as RO, synth  # noqa: E402
from instancefusion_b200.klg import KlgReader, KlgWriter  # noqa: E402


def maps_from_depth(depth_mm: torch.Tensor, K):
    """RGBA32F vertex / normal textures of a filtered depth image (camera frame), zeros where invalid."""
    H, W = depth_mm.shape
    z = depth_mm.to(torch.float32) / 1000.0
    v, u = torch.meshgrid(torch.arange(H, device=z.device, dtype=torch.float32), torch.arange(W, device=z.device, dtype=torch.float32),
                          indexing="ij")
    vm = torch.stack([(u - K.cx) * z / K.fx, (v - K.cy) * z / K.fy, z], -1)
    valid = z > 0
    dx = torch.zeros_like(vm); dy = torch.zeros_like(vm)
    dx[:, :-1] = vm[:, 1:] - vm[:, :-1]
    dy[:-1, :] = vm[1:, :] - vm[:-1, :]
    n = torch.cross(dx, dy, dim=-1)
    ok = valid.clone()
    ok[:, :-1] &= valid[:, 1:]; ok[:-1, :] &= valid[1:, :]
    ok[:, -1] = False; ok[-1, :] = False
    nn = n.norm(dim=-1, keepdim=True)
    ok &= nn[..., 0] > 0
    n = torch.where(ok[..., None], n / nn.clamp_min(1e-20), torch.zeros_like(n))
    one = torch.ones(H, W, 1, device=z.device)
    vm4 = torch.where(ok[..., None], torch.cat([vm, one], -1), torch.zeros(H, W, 4, device=z.device))
    nm4 = torch.where(ok[..., None], torch.cat([n, one], -1), torch.zeros(H, W, 4, device=z.device))
    return vm4.contiguous(), nm4.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--klg", default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="write a synthetic log of N frames first (and score against its ground truth)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--max-depth", type=float, default=20.0, help="filterDepth's maxD (ElasticFusion's depthCutoff)")
    ap.add_argument("--icp-weight", type=float, default=10.0)
    args = ap.parse_args()
    W, H = args.width, args.height
    K = synth.Intrinsics.kinect(W, H)
    gt = None
    path = args.klg
    if args.synthetic > 0:
        path = path or "/tmp/synth.klg"
        traj = synth.trajectory(args.synthetic, seed=2024)
        with KlgWriter(path, W, H, compress_depth=True, jpeg_quality=95) as wr:
            for k in range(args.synthetic):
                f = synth.render(traj[k], K, frame_id=k, device="cuda")
                d = f["depth"].view(torch.int16).cpu().numpy().view(np.uint16)
                wr.write(int(k * 1e6 / 30), d, f["rgba"][..., :3].cpu().numpy())
        gt = traj.numpy()
        print(f"wrote {path}: {args.synthetic} frames, {os.path.getsize(path) / 1e6:.1f} MB")
    assert path, "--klg or --synthetic"

    rd = KlgReader(path, W, H)
    trk = ef.RGBDOdometry(W, H, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)
    pose = (gt[0] if gt is not None else np.eye(4)).astype(np.float32)
    poses = [pose.copy()]
    prev = None
    t_track = 0.0
    for k, (ts, depth, rgb) in enumerate(rd):
        rgba = np.concatenate([rgb, np.full((H, W, 1), 255, np.uint8)], -1)
        d_raw = torch.from_numpy(depth.view(np.int16)).cuda().view(torch.uint16)
        d_rgba = torch.from_numpy(rgba).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        trk.initICPRaw(d_raw, args.max_depth, 20.0)            # filterDepth + initICP
        if prev is not None:
            vm4, nm4, p_rgba = prev
            trk.initICPModel(vm4, nm4, 20.0, pose)             # "model" = previous frame at its estimated pose
            trk.initRGBModel(p_rgba)
            trk.initRGB(d_rgba)
            t, R = trk.getIncrementalTransformation(pose[:3, 3].copy(), pose[:3, :3].copy(), False, args.icp_weight, True, False, False)
            pose = np.eye(4, dtype=np.float32)
            pose[:3, :3], pose[:3, 3] = R, t
            poses.append(pose.copy())
        t_track += time.perf_counter() - t0
        filt = torch.from_numpy(trk.buffer("filt_depth", 0).view(np.int16)).cuda().view(torch.uint16)
        vm4, nm4 = maps_from_depth(filt.to(torch.int32), K)
        prev = (vm4, nm4, d_rgba)
    n = len(poses)
    print(f"tracked {n - 1} frames, {(n - 1) / t_track:.0f} frames/s incl. host->device copies of the decoded frames "
          f"(log decoding and the torch model-map construction not counted)")
    if gt is not None:
        est = np.stack(poses)
        err = np.linalg.norm(est[:, :3, 3] - gt[:n, :3, 3], axis=1)
        path_len = np.linalg.norm(np.diff(gt[:n, :3, 3], axis=0), axis=1).sum()
        print(f"frame-to-frame drift vs ground truth: ATE rmse {np.sqrt((err ** 2).mean()) * 1000:.2f} mm, end {err[-1] * 1000:.2f} mm "
              f"over a {path_len:.2f} m path")
    trk.close()
    rd.close()


if __name__ == "__main__":
    main()
