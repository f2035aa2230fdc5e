#!/usr/bin/env python
"""Closed-loop tracking without OpenGL: surfel map -> CUDA model prediction (ef_op_splat_predict) at the LAST ESTIMATED pose ->
fill-in -> frame-to-model tracker (ef_track_frame_to_model) -> next pose, all buffers device-resident.

There is no surfel fusion in this repository (out of scope), so the map is a keyframe map: it is (re)seeded from
the current frame at its ESTIMATED pose every --keyframe frames, the way GlobalModel::initialise seeds the map from frame 1.
Every pose depends on the previous one, errors accumulate: the absolute trajectory error against the synthetic ground truth
is what the bench's open-loop protocol cannot show.

    python tools/closed_loop.py --frames 300 --keyframe 30
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
from instancefusion_b200.predict import ModelPredictor  # noqa: E402


def seed_surfels(pose, frame, K, t):
    s = synth.surfels_from_frame(pose, frame["vmap"].cpu().numpy(), frame["nmap"].cpu().numpy(), frame["rgba"].cpu().numpy(), K, time=t)
    return torch.from_numpy(s).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--keyframe", type=int, default=30)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--icp-weight", type=float, default=10.0)
    ap.add_argument("--fill-in", choices=["auto", "never", "always"], default="auto")
    args = ap.parse_args()
    K = synth.Intrinsics.kinect(args.width, args.height)
    gt = synth.trajectory(args.frames, seed=args.seed).numpy()
    frames = [synth.render(torch.from_numpy(gt[k]), K, seed=args.seed, frame_id=k, device="cuda") for k in range(args.frames)]
    trk = ef.RGBDOdometry(args.width, args.height, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)
    pred = ModelPredictor(args.width, args.height, K.cx, K.cy, K.fx, K.fy)
    pose = gt[0].astype(np.float32).copy()
    est = [pose.copy()]
    surfels = seed_surfels(pose, frames[0], K, 1)
    torch.cuda.synchronize()
    t_pred = t_track = 0.0
    e0, e1, e2, ep = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    t_splat = 0.0
    wall0 = time.perf_counter()
    seeds = 0.0
    filled = 0
    for k in range(1, args.frames):
        f = frames[k]
        e0.record()
        pred.predict(surfels, pose, time=k + 1, maxTime=k + 1, timeDelta=10 ** 6, maxDepth=20.0, confThreshold=9.0)
        ep.record()
        img, v, n = pred.image, pred.vertex, pred.normal
        # ElasticFusion.cpp:336-346: the filled maps are used only when the predicted colour image is not "dense enough"
        # (<= 75 % of its pixels non-black, :252-267; the reference looks at a downsampled copy)
        dense = float((img[::8, ::8, :3] > 0).all(-1).float().mean()) if args.fill_in == "auto" else 0.0
        if args.fill_in == "always" or (args.fill_in == "auto" and dense <= 0.75):
            img, v, n = pred.fill_in(f["depth"], f["rgba"], passthrough=False)
            filled += 1
        e1.record()
        with torch.cuda.stream(torch.cuda.ExternalStream(trk.stream)):
            torch.cuda.current_stream().wait_event(e1)
        t, R = trk.trackFrameToModel(v, n, img, f["depth"], f["rgba"], 20.0, pose, False, args.icp_weight, True, False, False)
        e2.record(torch.cuda.ExternalStream(trk.stream))
        e2.synchronize()
        t_pred += e0.elapsed_time(e1)
        t_splat += e0.elapsed_time(ep)
        t_track += e1.elapsed_time(e2)
        pose = pose.copy()
        pose[:3, :3], pose[:3, 3] = R, t
        est.append(pose.copy())
        if k % args.keyframe == 0:
            s0 = time.perf_counter()
            surfels = seed_surfels(pose, f, K, k + 1)  # host-side helper (numpy): not part of the timed pipeline
            seeds += time.perf_counter() - s0
    wall = time.perf_counter() - wall0 - seeds
    est = np.stack(est)
    err = np.linalg.norm(est[:, :3, 3] - gt[:, :3, 3], axis=1)
    step = np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1)
    n = args.frames - 1
    print(f"closed loop, {args.frames} frames {args.width}x{args.height}, keyframe map every {args.keyframe} frames, path length {step.sum():.2f} m")
    print(f"  absolute trajectory error: rmse {np.sqrt((err ** 2).mean()) * 1e3:.2f} mm, final {err[-1] * 1e3:.2f} mm, max {err.max() * 1e3:.2f} mm")
    print(f"  fill-in used on {filled} of {n} frames ({args.fill_in})")
    print(f"  device time per frame: prediction {t_splat / n * 1e3:.0f} us (+ density check / fill-in policy on the host: {(t_pred - t_splat) / n * 1e3:.0f} us), "
          f"builders + tracker {t_track / n * 1e3:.0f} us; "
          f"{n / wall:.0f} frames/s wall (python loop, keyframe seeding excluded)")
    trk.close()


if __name__ == "__main__":
    main()
