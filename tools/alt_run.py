"""A few batched launches (ef_track_frames_to_model_batch, k sequences) for ncu: python tools/alt_run.py [k] [launches]"""
import sys

import numpy as np

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w, h = 640, 480
K = synth.Intrinsics.kinect(w, h)
seqs = []
for seed in (2024, 7, 11, 99)[:k]:
    poses = synth.trajectory(3, seed=seed)
    seqs.append((poses.numpy().astype(np.float32), [synth.render(poses[i], K, seed=seed, frame_id=i, device="cuda") for i in range(3)]))
hs = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
bt = RO.BatchTracker(hs)
for i in range(n):
    f = 1 + i % 2
    fr = [(s[1][f - 1]["vmap"], s[1][f - 1]["nmap"], s[1][f - 1]["rgba"], s[1][f]["depth"], s[1][f]["rgba"]) for s in seqs]
    bt.track(fr, [s[0][f - 1] for s in seqs], 20.0, False, 10.0, True, False, False)
print("done")
