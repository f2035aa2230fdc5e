import sys, time
import numpy as np, torch
import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth
w, h = 640, 480
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = synth.Intrinsics.kinect(w, h)
sms = 148
seqs = []
for seed in (2024, 7, 11, 99)[:k]:
    poses = synth.trajectory(6, seed=seed)
    frames = [synth.render(poses[i], K, seed=seed, frame_id=i, device="cuda") for i in range(6)]
    seqs.append((poses.numpy().astype(np.float32), frames))
MODES = {
    "joint": (False, 10.0, True, False, False),
    "joint_so3": (False, 10.0, True, False, True),
    "icp_only": (False, 100.0, True, False, False),
    "rgb_only": (True, 10.0, True, False, False),
    "fast_nopyr": (False, 10.0, False, True, True),
}
P = lambda *a: print(*a, file=sys.stderr, flush=True)
single = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
for t in single:
    t.set_option(RO.EF_OPT_GRID_CTAS, sms - k + 1)
batched = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in range(k)]
bt = RO.BatchTracker(batched)
for g in range(k):
    batched[g].initFirstRGB(seqs[g][1][0]["rgba"])
    single[g].initFirstRGB(seqs[g][1][0]["rgba"])
for name, m in MODES.items():
    for f in range(1, 4):
        fr = [(s[1][f - 1]["vmap"], s[1][f - 1]["nmap"], s[1][f - 1]["rgba"], s[1][f]["depth"], s[1][f]["rgba"]) for s in seqs]
        ps = [s[0][f - 1] for s in seqs]
        P("single", name, f)
        want = [single[g].trackFrameToModel(*fr[g], 20.0, ps[g], *m) for g in range(k)]
        P("batched", name, f)
        got = bt.track(fr, ps, 20.0, *m)
        same = all(np.array_equal(got[g][0], want[g][0]) and np.array_equal(got[g][1], want[g][1]) for g in range(k))
        P("  same bits" if same else "  DIFFERENT", [float(np.abs(got[g][0] - want[g][0]).max()) for g in range(k)])
