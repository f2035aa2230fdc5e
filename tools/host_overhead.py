"""Where does a tracked frame's wall time go on the host side?  launch (enqueue of the builders + the tracker kernel)
vs finish (wait for the pinned result).  Developer tool; run on a GPU box."""
import sys, time, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO

class A: pass
args = A(); args.width, args.height, args.frames = 640, 480, 40
dev = torch.device("cuda:0")
K, poses, depth, rgba, vmap, nmap = bench.render_sequence(args, 2024, dev)
posef = poses.astype(np.float32)
tr = ef.RGBDOdometry(640, 480, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)
def one(i):
    k = 1 + (i % 39)
    t0 = time.perf_counter()
    tr.trackFrameToModelLaunch(vmap[k-1], nmap[k-1], rgba[k-1], depth[k], rgba[k], 20.0, posef[k-1], False, 10.0, True, False, False)
    t1 = time.perf_counter()
    tr.finish()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1
for i in range(30): one(i)
L, Fi = [], []
for i in range(300):
    a, b = one(i); L.append(a); Fi.append(b)
print(f"launch {np.median(L)*1e6:.1f} us  finish {np.median(Fi)*1e6:.1f} us  total {(np.median(L)+np.median(Fi))*1e6:.1f} us")
for aux in (0, 1):
    tr.set_option(RO.EF_OPT_AUX_STREAMS, aux)
    for i in range(30): one(i)
    L, Fi = [], []
    for i in range(300):
        a, b = one(i); L.append(a); Fi.append(b)
    print(f"aux_streams={aux}: launch {np.median(L)*1e6:.1f} us  finish {np.median(Fi)*1e6:.1f} us  total {(np.median(L)+np.median(Fi))*1e6:.1f} us")
tr.close()
