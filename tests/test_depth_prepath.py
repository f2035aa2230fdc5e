"""The step before the tracker: ElasticFusion::filterDepth / metriciseDepth, GLSL in the reference
(Shaders/depth_bilateral.frag, depth_metric.frag), CUDA here.  The CPU tests pin the restatement's properties; the GPU
tests compare the CUDA kernels with it.  Bar: metres bit-exact; filtered millimetres equal except where the two exp
implementations (libm / CUDA, <= 2 ulp) move a quotient across a .5 rounding tie (<= 1 mm, < 0.1 % of the pixels)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util


def _raw_depth(w, h, seed=5):
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    rng = np.random.default_rng(seed)
    d = f1["depth"].astype(np.int32)
    d = d + np.rint(rng.normal(0.0, 4.0, d.shape)).astype(np.int32)          # sensor noise, a few mm
    d = np.clip(d, 0, 65535).astype(np.uint16)
    d = util.punch_holes(d)
    d[40:60, 100:140] = 120                                                   # nearer than 300 mm: gated out
    d[200:220, 300:340] = 9000                                                # beyond maxD: gated out
    return d


def test_oracle_bilateral_properties():
    d = _raw_depth(160, 120)
    out = O.depth_bilateral(d, 4.0)
    gate = (d > 4000) | (d < 300)
    assert np.all(out[gate] == 0)                                             # depth_bilateral.frag:36-39
    assert np.all(out[~gate] > 0)
    # a constant image is a fixed point, borders included (clipped windows renormalise)
    c = np.full((48, 64), 1234, np.uint16)
    assert np.array_equal(O.depth_bilateral(c, 4.0), c)
    # the filter smooths: noise shrinks on a flat wall
    rng = np.random.default_rng(1)
    wall = (2000 + np.rint(rng.normal(0, 5, (64, 64)))).astype(np.uint16)
    f = O.depth_bilateral(wall, 4.0).astype(np.float64)
    assert f[8:-8, 8:-8].std() < 0.5 * wall[8:-8, 8:-8].astype(np.float64).std()
    # ... and keeps a 300 mm depth edge (the range kernel sigma is 30 mm)
    step = np.full((32, 64), 1000, np.uint16)
    step[:, 32:] = 1300
    fs = O.depth_bilateral(step, 4.0)
    assert np.array_equal(fs, step)


def test_oracle_metric():
    d = np.array([[0, 299, 300, 1234, 4000, 4001, 65535]], np.uint16)
    m = O.depth_metric(d, 4.0)
    assert np.array_equal(m, np.array([[0, 0, np.float32(300) / np.float32(1000), np.float32(1234) / np.float32(1000), 4.0, 0, 0]], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(640, 480), (1280, 720), (200, 152), (33, 17)])
def test_cuda_depth_prepath_matches_oracle(size):
    from instancefusion_b200 import ops
    w, h = size
    d = _raw_depth(max(w, 96), max(h, 64))[:h, :w].copy()
    for max_d in (4.0, 20.0):
        got, want = ops.depthBilateral(d, max_d), O.depth_bilateral(d, max_d)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, (size, max_d, int(diff.max()), float((diff != 0).mean()))
        assert np.array_equal(ops.depthMetric(d, max_d), O.depth_metric(d, max_d))


@pytest.mark.gpu
def test_raw_depth_entry_equals_filter_then_init():
    """ef_init_icp_depth_raw == ef_op_depth_bilateral followed by ef_init_icp_depth (device and host inputs)"""
    import torch
    import instancefusion_b200 as ef
    from instancefusion_b200 import ops
    w, h = 640, 480
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    raw = _raw_depth(w, h)
    a = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    b = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy)
    try:
        filt = ops.depthBilateral(raw, 4.0)
        a.initICP(filt, 20.0)
        b.initICPRaw(raw, 4.0, 20.0)
        assert np.array_equal(b.buffer("filt_depth", 0), filt)
        for lvl in range(3):
            assert np.array_equal(a.buffer("depth_tmp", lvl), b.buffer("depth_tmp", lvl))
            for name in ("vmap_curr", "nmap_curr"):
                assert np.array_equal(a.buffer(name, lvl), b.buffer(name, lvl), equal_nan=True), (name, lvl)
        b.initICPRaw(torch.from_numpy(raw.view(np.int16)).cuda().view(torch.uint16), 4.0, 20.0)
        assert np.array_equal(b.buffer("filt_depth", 0), filt)
    finally:
        a.close()
        b.close()
