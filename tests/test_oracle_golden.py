"""CPU tests (no GPU): pin the C restatement oracle/ef_oracle.c against golden vectors produced by the
REFERENCE's own CUDA kernels on a B200 (tests/golden/ref_cuda_160x120.npz, made by tests/golden/make_golden.py).

Documented deviations of a CPU restatement (oracle/ef_oracle.h): IEEE division / sqrt instead of the GPU's
approximate ones => float outputs within a few ulp, integer-valued outputs may differ by one LSB where the
quotient sits on an integer boundary, a borderline correspondence may flip."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cuda_160x120.npz"))
H, W = 120, 160
fx, fy, cx, cy = [np.float32(v) for v in G["K"]]


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_depth_pyramid_u16():
    got = O.pyr_down_u16(G["depth1"])
    d = np.abs(got.astype(np.int64) - G["pyr_down_u16"].astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 0.02


def test_vertex_and_normal_maps():
    v = O.create_vmap(G["depth1"], fx, fy, cx, cy, 20.0)
    assert util.masked_map_compare(v, G["vmap_curr"], H) <= 1
    n = O.create_nmap(G["vmap_curr"])
    assert np.array_equal(np.isnan(n[:H]), np.isnan(G["nmap_curr"][:H]))
    ok = ~np.isnan(G["nmap_curr"][:H])
    for c in range(3):
        assert np.allclose(n[c * H:(c + 1) * H][ok], G["nmap_curr"][c * H:(c + 1) * H][ok], atol=2e-6)


def test_copy_resize_transform():
    v, n = O.copy_maps(G["vmap0"], G["nmap0"])
    assert np.array_equal(v, G["copy_v"], equal_nan=True) and np.array_equal(n, G["copy_n"], equal_nan=True)
    assert util.masked_map_compare(O.resize_map(v, False), G["resize_v"], H // 2) == 0
    assert util.masked_map_compare(O.resize_map(n, True), G["resize_n"], H // 2) <= 4
    R = G["pose0"][:3, :3].astype(np.float32)
    t = G["pose0"][:3, 3].astype(np.float32)
    tv, tn = O.transform_maps(v, n, R, t)
    # with the FMA order of the reference binary restated (oracle/ef_oracle.c f3_dot) this is bit-exact
    assert util.masked_map_compare(tv, G["transform_v"], H) == 0
    assert util.masked_map_compare(tn, G["transform_n"], H) == 0


def test_rgbd_pyramids():
    d = O.vertices_to_depth(G["vmap0"], 6.0)
    assert np.array_equal(d, G["depth_f32"], equal_nan=True)
    d2 = O.pyr_down_gauss_f32(d)
    assert np.array_equal(np.isnan(d2), np.isnan(G["pyr_down_gauss_f32"]))
    assert int(util.ulp_diff(d2, G["pyr_down_gauss_f32"]).max()) <= 4
    assert np.array_equal(O.bgr_to_intensity(G["rgba0"]), G["intensity0"])
    i1 = O.bgr_to_intensity(G["rgba1"])
    assert np.array_equal(i1, G["intensity1"])
    p = O.pyr_down_gauss_u8(i1)
    dd = np.abs(p.astype(np.int64) - G["pyr_down_gauss_u8"].astype(np.int64))
    assert dd.max() <= 1 and (dd > 0).mean() < 0.02
    dx, dy = O.derivative_images(i1)
    assert np.array_equal(dx, G["dIdx"]) and np.array_equal(dy, G["dIdy"])
    cl = O.project_point_cloud(d, fx, fy, cx, cy, 0)
    assert int(util.ulp_diff(cl, G["cloud"]).max()) <= 1


def test_icp_step():
    R = G["pose0"][:3, :3].astype(np.float32)
    t = G["pose0"][:3, 3].astype(np.float32)
    Rinv = np.linalg.inv(R.astype(np.float64)).astype(np.float32)
    ang = float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0)))
    A, b, res = O.icp_step(R, t, G["vmap_curr"], G["nmap_curr"], Rinv, t, fx, fy, cx, cy, G["transform_v"], G["transform_n"], 0.10, ang)
    assert G["icp_res"][1] > 1000
    assert abs(res[1] - G["icp_res"][1]) <= max(3.0, 5e-4 * G["icp_res"][1])
    assert rel_err(A, G["icp_A"]) < 1e-4 and rel_err(b, G["icp_b"]) < 1e-3
    assert abs(res[0] - G["icp_res"][0]) <= 1e-3 * G["icp_res"][0]


def test_rgb_residual_and_step():
    cor, sig, cnt = O.rgb_residual(64.0, G["dIdx"], G["dIdy"], G["depth_f32"], G["depth_f32"], G["intensity0"], G["intensity1"], 0.07,
                                   G["kt"], G["krkinv"])
    gsig, gcnt = [int(v) for v in G["rgbres_sigma_count"]]
    assert gcnt > 300
    valid = cor["valid"] != 0
    flips = int((valid != G["rgbres_valid"]).sum())
    assert flips <= max(2, int(2e-3 * gcnt)), flips
    both = valid & G["rgbres_valid"]
    same = (cor["zero_x"][both] == G["rgbres_zero_x"][both]) & (cor["zero_y"][both] == G["rgbres_zero_y"][both])
    assert (~same).sum() <= max(2, int(2e-3 * gcnt))
    assert abs(cnt - gcnt) <= max(2, int(2e-3 * gcnt))
    # rgbStep on the GOLDEN correspondences
    rec = np.zeros((H, W), O.DATA_TERM)
    rec["valid"] = G["rgbres_valid"]
    rec["zero_x"], rec["zero_y"], rec["diff"] = G["rgbres_zero_x"], G["rgbres_zero_y"], G["rgbres_diff"]
    yy, xx = np.mgrid[0:H, 0:W]
    rec["one_x"], rec["one_y"] = xx, yy
    for name, sigma in (("w", float(G["rgb_sigma_w"])), ("unit", -1.0)):
        A, b = O.rgb_step(rec, sigma, G["cloud"], fx, fy, G["dIdx"], G["dIdy"], 0.125)
        assert rel_err(A, G["rgb_A_" + name]) < 1e-4 and rel_err(b, G["rgb_b_" + name]) < 1e-3, name


def test_so3_step():
    A, b, res = O.so3_step(G["intensity0"], G["intensity1"], G["so3_H"], G["so3_kinv"], G["so3_krlr"])
    assert res[1] == G["so3_res"][1]
    assert rel_err(A, G["so3_A"]) < 1e-4 and rel_err(b, G["so3_b"]) < 1e-3 and abs(res[0] - G["so3_res"][0]) <= 1e-4 * G["so3_res"][0]


@pytest.mark.parametrize("name,args", [("icp_nopyr", (False, 100.0, False, False, False)), ("joint", (False, 10.0, True, False, False)),
                                       ("joint_so3", (False, 10.0, True, False, True))])
def test_full_tracker(name, args):
    pose0f = G["pose0"].astype(np.float32)
    tr = O.OracleTracker(W, H, cx, cy, fx, fy)
    tr.init_first_rgb(G["rgba0"])
    tr.init_icp_model(G["vmap0"], G["nmap0"], 20.0, pose0f)
    tr.init_rgb_model(G["rgba0"])
    tr.init_icp_depth(G["depth1"], 20.0)
    tr.init_rgb(G["rgba1"])
    t, R, st = tr.get_incremental_transformation(pose0f[:3, 3], pose0f[:3, :3], *args)
    tr.close()
    assert st["se3_iterations"] + [st["so3_iterations"]] == [int(v) for v in G["track_%s_iters" % name]]
    assert float(np.abs(t - G["track_%s_t" % name]).max()) <= 5e-5
    assert util.rot_err(R, G["track_%s_R" % name]) <= 5e-5
    gst = G["track_%s_stats" % name]
    assert st["last_icp_count"] == pytest.approx(gst[1], rel=5e-3)
    assert st["last_rgb_count"] == pytest.approx(gst[3], rel=5e-3, abs=3)
    assert st["last_so3_count"] == gst[5]
    assert np.linalg.norm(st["last_A"] - G["track_%s_A" % name]) <= 2e-3 * np.linalg.norm(G["track_%s_A" % name])
    # and it tracks: the estimate is closer to the ground truth than the prior
    prior = np.linalg.norm(G["pose0"][:3, 3] - G["pose1"][:3, 3])
    assert np.linalg.norm(t - G["pose1"][:3, 3]) < prior
