"""The step around the tracker: model prediction (IndexMap::combinedPredict, an OpenGL point-sprite pass in
the reference) and FillIn as CUDA operators, against the CPU restatement oracle/ef_oracle.c (which states the rasterisation
rules GL leaves open -- parity against a GL driver is unpinned) and against the ray-cast ground truth of the synthetic
scene.  CPU tests cover the restatement itself."""
import numpy as np
import pytest

from instancefusion_b200 import synth
from oracle import oracle as O
from tests import util

W, H = 320, 240


def _scene(w=W, h=H, stride=12):
    K, pose0, pose1, f0, f1 = util.frame_pair(w, h)
    surfels = synth.surfels_from_frame(pose0, f0["vmap"], f0["nmap"], f0["rgba"], K, time=5, stride_floats=stride)
    return K, pose0, pose1, f0, f1, surfels


ARGS = dict(max_depth=20.0, conf_threshold=9.0, time=6, max_time=6, time_delta=200)


def _oracle(surfels, pose, K, **kw):
    a = dict(ARGS, **kw)
    return O.splat_predict(surfels, pose, K.cx, K.cy, K.fx, K.fy, K.height, K.width, a["max_depth"], a["conf_threshold"], a["time"], a["max_time"],
                           a["time_delta"])


def test_restatement_predicts_the_ray_cast_model():
    """known answer: surfels seeded from frame 0, drawn from pose 1, must reproduce the scene ray-cast from pose 1"""
    K, pose0, pose1, f0, f1, surfels = _scene()
    img, v, n, tm = _oracle(surfels, pose1, K)
    truth = f1["vmap"][..., 2]
    both = (v[..., 2] > 0) & (truth > 0)
    assert both.mean() > 0.85
    err = np.abs(v[..., 2] - truth)[both]
    assert np.median(err) < 2e-3 and np.percentile(err, 90) < 2e-2  # depth noise of frame 0 + disc approximation at edges
    nt = f1["nmap"][..., :3]
    good = both & np.isfinite(nt[..., 0])
    cosang = np.abs((n[..., :3] * nt).sum(-1))[good]
    assert np.median(cosang) > 0.99
    assert (tm[both] == 5).all() and (img[..., 3][both] == 255).all()
    # the vertex lies on the pixel's viewing ray through the pixel centre (combo_splat.frag:61)
    ys, xs = np.nonzero(both)
    assert np.allclose(v[ys, xs, 0], (xs + 0.5 - K.cx) * v[ys, xs, 2] / K.fx, rtol=1e-5, atol=1e-6)


def test_restatement_cull_rules():
    K, pose0, pose1, f0, f1, surfels = _scene()
    empty = lambda out: not out[1].any() and not out[2].any() and not out[0].any() and not out[3].any()
    assert empty(_oracle(surfels[:0], pose1, K))
    assert empty(_oracle(surfels, pose1, K, conf_threshold=11.0))          # splat.vert:51 confidence
    assert empty(_oracle(surfels, pose1, K, time=500, time_delta=200))     # too old
    assert empty(_oracle(surfels, pose1, K, max_time=4))                   # from the future
    assert empty(_oracle(surfels, pose1, K, max_depth=0.2))                # beyond the depth range
    behind = pose1.copy()
    behind[:3, :3] = behind[:3, :3] @ np.diag([-1.0, 1.0, -1.0])           # look the other way: everything has z < 0 or leaves the view
    out = _oracle(surfels, behind, K)
    assert (out[1][..., 2] >= 0).all()


def test_restatement_fill_in():
    K, pose0, pose1, f0, f1, surfels = _scene()
    _, v, n, _ = _oracle(surfels, pose1, K)
    v[40:80, 50:120] = 0
    fv = O.fill_vertex(v, f1["depth"], K.cx, K.cy, K.fx, K.fy)
    fn = O.fill_normal(v, f1["depth"], K.cx, K.cy, K.fx, K.fy)
    hole = v[..., 2] == 0
    assert np.array_equal(fv[~hole], v[~hole]) and np.array_equal(fn[~hole], v[~hole])
    z = f1["depth"].astype(np.float32) / 1000.0
    assert np.allclose(fv[..., 2][hole], z[hole]) and (fv[..., 3][hole] == 1).all()
    ys, xs = np.nonzero(hole)
    assert np.allclose(fv[ys, xs, 0], (xs - K.cx) * z[ys, xs] / K.fx, rtol=1e-5, atol=1e-6)  # fill_vertex.frag: integer pixel coordinates
    allv = O.fill_vertex(v, f1["depth"], K.cx, K.cy, K.fx, K.fy, passthrough=True)
    assert np.allclose(allv[..., 2], z)
    img = f0["rgba"].copy()
    img[10:20, 10:20, :3] = 0
    out = O.fill_rgb(img, f1["rgba"])
    assert np.array_equal(out[10:20, 10:20], f1["rgba"][10:20, 10:20]) and np.array_equal(out[30:], img[30:])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("stride", [12, 64])
@pytest.mark.parametrize("size", [(320, 240), (640, 480)])
def test_splat_matches_restatement_bit_for_bit(size, stride):
    from instancefusion_b200 import ops
    w, h = size
    K, pose0, pose1, f0, f1, surfels = _scene(w, h, stride)
    rng = np.random.default_rng(3)
    surfels = surfels[rng.permutation(len(surfels))]     # submission order != raster order: exercises the depth-test tie-break
    surfels[::7, 3] = 5.0                                 # some below the confidence threshold
    surfels[::11, 7] = 900.0                              # some newer than maxTime
    surfels[::13, 11] *= 6.0                              # some large discs (overlapping sprites)
    for pose in (pose1, pose0):
        got = ops.splatPredict(surfels, pose, K.cx, K.cy, K.fx, K.fy, h, w, ARGS["max_depth"], ARGS["conf_threshold"], ARGS["time"],
                               ARGS["max_time"], ARGS["time_delta"])
        want = _oracle(surfels, pose, K)
        for g, e, name in zip(got, want, ("image", "vertex", "normal", "time")):
            assert np.array_equal(g, e), name
    got = ops.splatPredict(surfels[:0], pose1, K.cx, K.cy, K.fx, K.fy, h, w, 20.0, 9.0, 6, 6, 200)
    assert not got[1].any() and not got[0].any()


@pytest.mark.gpu
def test_fill_in_matches_restatement_bit_for_bit():
    from instancefusion_b200 import ops
    K, pose0, pose1, f0, f1, surfels = _scene()
    _, v, n, _ = _oracle(surfels, pose1, K)
    v[40:80, 50:120] = 0
    depth = util.punch_holes(f1["depth"])
    for pt in (False, True):
        assert np.array_equal(ops.fillVertex(v, depth, K.cx, K.cy, K.fx, K.fy, pt), O.fill_vertex(v, depth, K.cx, K.cy, K.fx, K.fy, pt))
        assert np.array_equal(ops.fillNormal(v, depth, K.cx, K.cy, K.fx, K.fy, pt), O.fill_normal(v, depth, K.cx, K.cy, K.fx, K.fy, pt),
                              equal_nan=True)
        img = f0["rgba"].copy()
        img[10:20, 10:20, :3] = 0
        assert np.array_equal(ops.fillImage(img, f1["rgba"], pt), O.fill_rgb(img, f1["rgba"], pt))


@pytest.mark.gpu
def test_tracking_against_the_predicted_model():
    """closed loop without GL: surfels -> CUDA prediction at the prior pose -> tracker; the pose it returns is as good as with
    the ray-cast model maps"""
    import instancefusion_b200 as ef
    from instancefusion_b200 import ops, rgbd_odometry as RO
    w, h = 640, 480
    K, pose0, pose1, f0, f1, surfels = _scene(w, h)
    img, v, n, _ = ops.splatPredict(surfels, pose0, K.cx, K.cy, K.fx, K.fy, h, w, 20.0, 9.0, 6, 6, 200)
    tr = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)
    try:
        p0 = pose0.astype(np.float32)
        t, R = tr.trackFrameToModel(v, n, img, f1["depth"], f1["rgba"], 20.0, p0, False, 10.0, True, False, False)
        assert np.linalg.norm(t - pose1[:3, 3]) < 3e-3 and util.rot_err(R, pose1[:3, :3]) < 2e-3
    finally:
        tr.close()


@pytest.mark.gpu
def test_closed_loop_over_a_trajectory_stays_on_track():
    """12 frames of the bench trajectory in closed loop, every buffer device-resident: map seeded from frame 0 -> CUDA prediction
    at the last ESTIMATED pose -> tracker -> next pose.  The pose error must stay at the millimetre level (the prediction samples
    the surface through fragment CENTRES, i + 0.5, while the tracker's own maps use integer pixel coordinates -- the reference's
    GL / CUDA conventions differ by that half pixel as well -- so the photometric term carries a bias of ~z / fx / 2)."""
    import torch
    import instancefusion_b200 as ef
    from instancefusion_b200 import rgbd_odometry as RO
    from instancefusion_b200.predict import ModelPredictor
    w, h, n = 640, 480, 12
    K = synth.Intrinsics.kinect(w, h)
    gt = synth.trajectory(n, seed=2024).numpy()
    frames = [synth.render(torch.from_numpy(gt[k]), K, seed=2024, frame_id=k, device="cuda") for k in range(n)]
    f0 = frames[0]
    surfels = torch.from_numpy(synth.surfels_from_frame(gt[0], f0["vmap"].cpu().numpy(), f0["nmap"].cpu().numpy(), f0["rgba"].cpu().numpy(), K)).cuda()
    trk = ef.RGBDOdometry(w, h, K.cx, K.cy, K.fx, K.fy, solve_mode=RO.EF_SOLVE_DEVICE)
    pred = ModelPredictor(w, h, K.cx, K.cy, K.fx, K.fy)
    try:
        pose = gt[0].astype(np.float32).copy()
        errs = []
        for k in range(1, n):
            img, v, nm, _ = pred.predict(surfels, pose, time=k + 1, timeDelta=10 ** 6, confThreshold=9.0)
            torch.cuda.current_stream().synchronize()
            t, R = trk.trackFrameToModel(v, nm, img, frames[k]["depth"], frames[k]["rgba"], 20.0, pose, False, 10.0, True, False, False)
            pose = pose.copy()
            pose[:3, :3], pose[:3, 3] = R, t
            errs.append(float(np.linalg.norm(t - gt[k][:3, 3])))
        assert max(errs) < 6e-3, errs
    finally:
        trk.close()


@pytest.mark.gpu
def test_instance_render_target():
    """InstanceFusion's fifth target of combo_splat.frag (location 4, :29 / :54): inst = decodeColor(colTime.y) of the surfel that
    wins the depth test -- the same surfel whose colour lands in `image`, so with instance colour := colour XOR a constant the two
    targets must be that XOR of each other pixel by pixel, and zero together where nothing was drawn."""
    import torch
    from instancefusion_b200.predict import ModelPredictor
    w, h = 320, 240
    K, pose0, pose1, f0, f1, surfels = _scene(w, h, 64)  # InstanceFusion's 256-byte vertex
    col = surfels[:, 4].astype(np.int64)
    surfels[:, 5] = (col ^ 0x00A5C3).astype(np.float32)  # 24-bit values are exact in float32
    pred = ModelPredictor(w, h, K.cx, K.cy, K.fx, K.fy)
    pred.predict(torch.from_numpy(surfels).cuda(), pose1, time=ARGS["time"], maxTime=ARGS["max_time"], timeDelta=ARGS["time_delta"],
                 maxDepth=ARGS["max_depth"], confThreshold=ARGS["conf_threshold"])
    torch.cuda.synchronize()
    img, inst = pred.image.cpu().numpy().astype(np.int64), pred.inst.cpu().numpy().astype(np.int64)
    drawn = img[..., 3] == 255
    assert drawn.mean() > 0.5
    rgb = (img[..., 0] << 16) | (img[..., 1] << 8) | img[..., 2]
    irgb = (inst[..., 0] << 16) | (inst[..., 1] << 8) | inst[..., 2]
    assert np.array_equal(irgb[drawn], rgb[drawn] ^ 0x00A5C3)
    assert np.array_equal(inst[..., 3] == 255, drawn) and not inst[~drawn].any()
