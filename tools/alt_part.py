"""P partitions x k alternating sequences: python tools/alt_part.py P k"""
import sys, time
import numpy as np, torch
import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth
P, k = int(sys.argv[1]), int(sys.argv[2])
w, h = 640, 480
n = 16
K = synth.Intrinsics.kinect(w, h)
sms = torch.cuda.get_device_properties(0).multi_processor_count
seqs = []
for seed in (2024, 7, 11, 99)[:k]:
    poses = synth.trajectory(n, seed=seed)
    frames = [synth.render(poses[i], K, seed=seed, frame_id=i, device="cuda") for i in range(n)]
    seqs.append((poses.numpy().astype(np.float32), frames))
m = (False, 10.0, True, False, False)
bts = []
for p in range(P):
    hs = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
    for g, t in enumerate(hs):
        t.initFirstRGB(seqs[g][1][0]["rgba"])
        if P > 1:
            t.set_option(RO.EF_OPT_GRID_CTAS, sms // P)
    bts.append(RO.BatchTracker(hs) if k > 1 else hs[0])

def batch(i):
    f = 1 + i % (n - 1)
    fr = [(s[1][f - 1]["vmap"], s[1][f - 1]["nmap"], s[1][f - 1]["rgba"], s[1][f]["depth"], s[1][f]["rgba"]) for s in seqs[:k]]
    return fr, [s[0][f - 1] for s in seqs[:k]]

def launch(b, i):
    fr, ps = batch(i)
    if k > 1:
        b.launch(fr, ps, 20.0, *m)
    else:
        b.trackFrameToModelLaunch(*fr[0], 20.0, ps[0], *m)

for rep in range(2):
    N = 300
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(P):
        launch(bts[i], i)
    for i in range(P, N):
        bts[i % P].finish()
        launch(bts[i % P], i)
    for i in range(P):
        bts[i].finish()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"P={P} partitions of {sms // P} SMs x k={k}: {k * N / dt:8.0f} frames/s ({dt / N * 1e3:.3f} ms per batch)", flush=True)
