"""k sequences per launch (ef_track_frames_to_model_batch, BASELINE.json configs[4] "k sequences per GPU"): the batched
persistent kernel gives every handle exactly the bits its own single launch gives, frame after frame, in every mode (a thread
group of the batched build has 4 warps, a single launch 8: each of its threads keeps the accumulators of the two virtual
threads it stands for, so the float sums are added in the single launch's order -- EF_TRACK_SETS in ef_track_kernel.cu)."""
import numpy as np
import pytest

import instancefusion_b200 as ef
from instancefusion_b200 import rgbd_odometry as RO
from instancefusion_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu

MODES = {
    "joint": dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=False),
    "joint_so3": dict(rgbOnly=False, icpWeight=10.0, pyramid=True, fastOdom=False, so3=True),
    "icp_only": dict(rgbOnly=False, icpWeight=100.0, pyramid=True, fastOdom=False, so3=False),
    "rgb_only": dict(rgbOnly=True, icpWeight=10.0, pyramid=True, fastOdom=False, so3=False),
    "fast_nopyr": dict(rgbOnly=False, icpWeight=10.0, pyramid=False, fastOdom=True, so3=True),
}


def _sequences(w, h, n, seeds):
    K = synth.Intrinsics.kinect(w, h)
    out = []
    for seed in seeds:
        poses = synth.trajectory(n, seed=seed)
        frames = [synth.render(poses[k], K, seed=seed, frame_id=k, device="cuda") for k in range(n)]
        out.append((poses.numpy().astype(np.float32), frames))
    return K, out


@pytest.mark.parametrize("size", [(640, 480), (320, 240), (322, 242)])
def test_batched_launch_equals_single_launches_bit_for_bit(size):
    w, h = size
    n = 5
    K, seqs = _sequences(w, h, n, seeds=(2024, 7))
    single = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in seqs]
    batched = [ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE) for _ in seqs]
    bt = RO.BatchTracker(batched)
    try:
        for g, (poses, frames) in enumerate(seqs):
            single[g].initFirstRGB(frames[0]["rgba"])
            batched[g].initFirstRGB(frames[0]["rgba"])
        for name, m in MODES.items():
            for k in range(1, n):  # consecutive frames: state (SO(3) image swap, last* fields) carried
                args = (20.0, m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
                fr = [(f[k - 1]["vmap"], f[k - 1]["nmap"], f[k - 1]["rgba"], f[k]["depth"], f[k]["rgba"]) for _, f in seqs]
                ps = [p[k - 1] for p, _ in seqs]
                want = [single[g].trackFrameToModel(*fr[g], 20.0, ps[g], *args[1:]) for g in range(len(seqs))]
                got = bt.track(fr, ps, *args)
                for g in range(len(seqs)):
                    assert np.array_equal(got[g][0], want[g][0]) and np.array_equal(got[g][1], want[g][1]), (name, k, g, got[g][0], want[g][0])
                    a, b = batched[g], single[g]
                    assert a.lastICPCount == b.lastICPCount and a.lastRGBCount == b.lastRGBCount and a.lastSO3Count == b.lastSO3Count, (name, k, g)
                    assert a.lastICPError == b.lastICPError and a.lastRGBError == b.lastRGBError, (name, k, g)
                    assert np.array_equal(a.lastA, b.lastA) and np.array_equal(a.lastb, b.lastb), (name, k, g)
                    assert a.se3_iterations == b.se3_iterations and a.so3_iterations == b.so3_iterations
        m = MODES["joint"]
        l0 = [t.launch_count for t in batched]
        bt.track(fr, ps, 20.0, m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
        assert [t.launch_count - x for t, x in zip(batched, l0)] == [2, 2] or w % 4 != 0
        # and the batched launch reproduces itself to the bit
        again = bt.track(fr, ps, 20.0, m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
        once_more = bt.track(fr, ps, 20.0, m["rgbOnly"], m["icpWeight"], m["pyramid"], m["fastOdom"], m["so3"])
        for g in range(len(seqs)):
            assert np.array_equal(again[g][0], once_more[g][0]) and np.array_equal(again[g][1], once_more[g][1])
    finally:
        for t in single + batched:
            t.close()


def test_batch_argument_checks():
    w, h = 320, 240
    K, seqs = _sequences(w, h, 2, seeds=(1, 2))
    a = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    b = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_HOST)
    c = ef.RGBDOdometry(2 * w, 2 * h, K.cx, K.cy, K.fx, K.fy)
    try:
        with pytest.raises(ValueError):
            RO.BatchTracker([a])
        fr = [(f[0]["vmap"], f[0]["nmap"], f[0]["rgba"], f[1]["depth"], f[1]["rgba"]) for _, f in seqs]
        ps = [p[0] for p, _ in seqs]
        for pair in ([a, a], [a, b], [a, c]):  # the same handle twice, a host-solve handle, another image size
            with pytest.raises((ef.EFError, ValueError)):
                RO.BatchTracker(pair).track(fr if pair[1] is not c else fr, ps, 20.0, False, 10.0, True, False, False)
    finally:
        for t in (a, b, c):
            t.close()
