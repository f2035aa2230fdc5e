"""Alternating batched kernel (EF_TRACK_ALT): bits vs single launches with the same number of workers, and throughput for
k = 2, 3, 4 sequences per launch.   python tools/alt_check.py [width height]"""
import sys
import time

import numpy as np
import torch

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
n = 16
K = synth.Intrinsics.kinect(w, h)
sms = torch.cuda.get_device_properties(0).multi_processor_count
seqs = []
for seed in (2024, 7, 11, 99):
    poses = synth.trajectory(n, seed=seed)
    frames = [synth.render(poses[k], K, seed=seed, frame_id=k, device="cuda") for k in range(n)]
    seqs.append((poses.numpy().astype(np.float32), frames))

MODES = {
    "joint": (False, 10.0, True, False, False),
    "joint_so3": (False, 10.0, True, False, True),
    "icp_only": (False, 100.0, True, False, False),
    "rgb_only": (True, 10.0, True, False, False),
    "fast_nopyr": (False, 10.0, False, True, True),
}

for k in (2, 3, 4):
    single = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
    batched = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
    for t in single:
        t.set_option(RO.EF_OPT_GRID_CTAS, sms - k)  # the same number of workers as the batched launch gives a sequence
    try:
        bt = RO.BatchTracker(batched)
        for g in range(k):
            single[g].initFirstRGB(seqs[g][1][0]["rgba"])
            batched[g].initFirstRGB(seqs[g][1][0]["rgba"])
        bad = 0
        for name, m in MODES.items():
            for f in range(1, 3):
                fr = [(s[1][f - 1]["vmap"], s[1][f - 1]["nmap"], s[1][f - 1]["rgba"], s[1][f]["depth"], s[1][f]["rgba"]) for s in seqs[:k]]
                ps = [s[0][f - 1] for s in seqs[:k]]
                want = [single[g].trackFrameToModel(*fr[g], 20.0, ps[g], *m) for g in range(k)]
                got = bt.track(fr, ps, 20.0, *m)
                for g in range(k):
                    same = np.array_equal(got[g][0], want[g][0]) and np.array_equal(got[g][1], want[g][1]) and np.array_equal(batched[g].lastA, single[g].lastA)
                    if not same:
                        bad += 1
                        print("MISMATCH", k, name, f, g, got[g][0], want[g][0])
        print(f"k={k}: bit comparison vs single launches on {sms - k} workers: {'OK' if bad == 0 else str(bad) + ' mismatches'}", flush=True)
        # throughput, blocking, one batch at a time
        m = MODES["joint"]

        def batch(i):
            f = 1 + i % (n - 1)
            fr = [(s[1][f - 1]["vmap"], s[1][f - 1]["nmap"], s[1][f - 1]["rgba"], s[1][f]["depth"], s[1][f]["rgba"]) for s in seqs[:k]]
            return fr, [s[0][f - 1] for s in seqs[:k]]

        for i in range(10):
            bt.track(*batch(i), 20.0, *m)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        N = 200
        for i in range(N):
            bt.track(*batch(i), 20.0, *m)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"k={k}: blocking {k * N / dt:8.0f} frames/s  ({dt / N * 1e3:.3f} ms per batch)", flush=True)
        # two batches in flight: a second set of handles is launched before the first is finished
        other = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
        for g in range(k):
            other[g].initFirstRGB(seqs[g][1][0]["rgba"])
        bts = [bt, RO.BatchTracker(other)]
        for i in range(4):
            bts[i % 2].launch(*batch(i), 20.0, *m)
            bts[i % 2].finish()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bts[0].launch(*batch(0), 20.0, *m)
        for i in range(1, N):
            bts[i % 2].launch(*batch(i), 20.0, *m)
            bts[(i - 1) % 2].finish()
        bts[(N - 1) % 2].finish()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"k={k}: two batches in flight {k * N / dt:8.0f} frames/s  ({dt / N * 1e3:.3f} ms per batch)", flush=True)
        for t in other:
            t.close()
    except Exception as e:  # noqa: BLE001
        print(f"k={k}: {e!r}", flush=True)
    finally:
        for t in single + batched:
            t.close()
