"""GPU parity of every Tier-2 operator (through the C ABI) against
  (a) the reference's OWN CUDA kernels (oracle/_ref/libef_ref.so, built unmodified from
      /root/reference by oracle/Makefile) -- bit-exact for integer outputs, <= 1 ulp for float maps,
      1e-4 norm-relative for the reduced normal equations, and
  (b) the CPU restatement oracle/ef_oracle.c -- a few ulp / one LSB (IEEE vs approximate GPU ops).
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

LEVEL_SHAPES = [(480, 640), (240, 320), (120, 160), (60, 80)]


@pytest.fixture(scope="module")
def ops():
    from instancefusion_b200 import ops as P
    return P


@pytest.fixture(scope="module")
def pair():
    return util.frame_pair(640, 480)


@pytest.fixture(scope="module")
def have_ref():
    if not O.ref_available():
        pytest.fail("oracle/_ref/libef_ref.so missing: parity against the reference CUDA cannot be checked")
    return True


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def int_mismatch_frac(a, b, tol=0):
    a, b = a.astype(np.int64), b.astype(np.int64)
    return float((np.abs(a - b) > tol).mean())


# ---------------------------------------------------------------------------------------------
def test_pyr_down_u16(ops, pair, have_ref):
    _, _, _, _, f1 = pair
    for depth in (f1["depth"], util.punch_holes(f1["depth"])):
        got = ops.pyrDown(depth)
        ref = O.pyr_down_u16(depth, impl="ref")
        assert np.array_equal(got, ref)
        cpu = O.pyr_down_u16(depth, impl="cpu")
        assert int_mismatch_frac(got, cpu, 1) == 0.0
        assert int_mismatch_frac(got, cpu, 0) < 0.02
        got2 = ops.pyrDown(got)
        assert np.array_equal(got2, O.pyr_down_u16(ref, impl="ref"))


def test_create_vmap_nmap(ops, pair, have_ref):
    K, _, _, _, f1 = pair
    depth = util.punch_holes(f1["depth"])
    for level in range(3):
        fx, fy, cx, cy = util.se3_level_params(K, level)
        rows = depth.shape[0]
        for cutoff in (20.0, 3.0):
            v = ops.createVMap(depth, fx, fy, cx, cy, cutoff)
            vr = O.create_vmap(depth, fx, fy, cx, cy, cutoff, impl="ref")
            assert util.masked_map_compare(v, vr, rows) == 0
            vc = O.create_vmap(depth, fx, fy, cx, cy, cutoff, impl="cpu")
            assert util.masked_map_compare(v, vc, rows) <= 1
            n = ops.createNMap(v)
            nr = O.create_nmap(vr, impl="ref")
            assert util.masked_map_compare(n, nr, rows) <= 1
            nc = O.create_nmap(vc, impl="cpu")
            gx, wx = n[:rows], nc[:rows]
            assert np.array_equal(np.isnan(gx), np.isnan(wx))
            ok = ~np.isnan(wx)
            for c in range(3):
                assert np.allclose(n[c * rows:(c + 1) * rows][ok], nc[c * rows:(c + 1) * rows][ok], atol=2e-6)
        depth = O.pyr_down_u16(depth, impl="ref")


def test_copy_resize_transform(ops, pair, have_ref):
    _, pose0, _, f0, _ = pair
    v4, n4 = util.holes_in_maps(f0["vmap"], f0["nmap"])
    v, n = ops.copyMaps(v4, n4)
    vr, nr = O.copy_maps(v4, n4, impl="ref")
    assert np.array_equal(v, vr, equal_nan=True) and np.array_equal(n, nr, equal_nan=True)
    vc, nc = O.copy_maps(v4, n4, impl="cpu")
    assert np.array_equal(v, vc, equal_nan=True) and np.array_equal(n, nc, equal_nan=True)

    R = pose0[:3, :3].astype(np.float32)
    t = pose0[:3, 3].astype(np.float32)
    rows = v.shape[0] // 3
    for level in range(3):
        tv, tn = ops.tranformMaps(v, n, R, t)
        tvr, tnr = O.transform_maps(vr, nr, R, t, impl="ref")
        assert util.masked_map_compare(tv, tvr, rows) == 0
        assert util.masked_map_compare(tn, tnr, rows) == 0
        tvc, tnc = O.transform_maps(vc, nc, R, t, impl="cpu")
        assert util.masked_map_compare(tv, tvc, rows) <= 2
        if level == 2:
            break
        v2, n2 = ops.resizeVMap(v), ops.resizeNMap(n)
        v2r, n2r = O.resize_map(vr, False, impl="ref"), O.resize_map(nr, True, impl="ref")
        assert util.masked_map_compare(v2, v2r, rows // 2) == 0
        assert util.masked_map_compare(n2, n2r, rows // 2) <= 1
        v2c, n2c = O.resize_map(vc, False, impl="cpu"), O.resize_map(nc, True, impl="cpu")
        assert util.masked_map_compare(v2, v2c, rows // 2) == 0
        assert util.masked_map_compare(n2, n2c, rows // 2) <= 4
        v, n, vr, nr, vc, nc, rows = v2, n2, v2r, n2r, v2c, n2c, rows // 2


def test_depth_pyramid_f32(ops, pair, have_ref):
    _, _, _, f0, _ = pair
    v4, _ = util.holes_in_maps(f0["vmap"], f0["nmap"])
    for cutoff in (6.0, 2.5):
        d = ops.verticesToDepth(v4, cutoff)
        dr = O.vertices_to_depth(v4, cutoff, impl="ref")
        assert np.array_equal(d, dr, equal_nan=True)
        assert np.array_equal(d, O.vertices_to_depth(v4, cutoff, impl="cpu"), equal_nan=True)
        for _ in range(2):
            d2 = ops.pyrDownGaussF(d)
            d2r = O.pyr_down_gauss_f32(dr, impl="ref")
            assert int(util.ulp_diff(d2, d2r).max()) == 0
            d2c = O.pyr_down_gauss_f32(dr, impl="cpu")
            assert np.array_equal(np.isnan(d2), np.isnan(d2c))
            assert int(util.ulp_diff(d2, d2c).max()) <= 4
            d, dr = d2, d2r


def test_intensity_pyramid_and_derivatives(ops, pair, have_ref):
    _, _, _, _, f1 = pair
    rgba = f1["rgba"].copy()
    rgba[100:120, 200:260] = 0  # "no data" region
    img = ops.imageBGRToIntensity(rgba)
    imr = O.bgr_to_intensity(rgba, impl="ref")
    assert np.array_equal(img, imr)
    assert np.array_equal(img, O.bgr_to_intensity(rgba, impl="cpu"))
    for level in range(3):
        dx, dy = ops.computeDerivativeImages(img)
        dxr, dyr = O.derivative_images(imr, impl="ref")
        assert np.array_equal(dx, dxr) and np.array_equal(dy, dyr)
        dxc, dyc = O.derivative_images(imr, impl="cpu")
        assert np.array_equal(dx, dxc) and np.array_equal(dy, dyc)
        if level == 2:
            break
        i2 = ops.pyrDownUcharGauss(img)
        i2r = O.pyr_down_gauss_u8(imr, impl="ref")
        assert np.array_equal(i2, i2r)
        i2c = O.pyr_down_gauss_u8(imr, impl="cpu")
        assert int_mismatch_frac(i2, i2c, 1) == 0.0 and int_mismatch_frac(i2, i2c, 0) < 0.02
        img, imr = i2, i2r
    # adversarial: random image with zeros (exercises the all-zero window -> NaN -> 0 conversion)
    rnd = util.random_image_u8(96, 128, zero_frac=0.6)
    assert np.array_equal(ops.pyrDownUcharGauss(rnd), O.pyr_down_gauss_u8(rnd, impl="ref"))
    dx, dy = ops.computeDerivativeImages(rnd)
    dxr, dyr = O.derivative_images(rnd, impl="ref")
    assert np.array_equal(dx, dxr) and np.array_equal(dy, dyr)


def test_project_point_cloud(ops, pair, have_ref):
    K, _, _, f0, _ = pair
    d = O.vertices_to_depth(f0["vmap"], 6.0, impl="cpu")
    for level in range(3):
        c = ops.projectToPointCloud(d, K.fx, K.fy, K.cx, K.cy, level)
        cr = O.project_point_cloud(d, K.fx, K.fy, K.cx, K.cy, level, impl="ref")
        assert np.array_equal(c, cr, equal_nan=True)
        cc = O.project_point_cloud(d, K.fx, K.fy, K.cx, K.cy, level, impl="cpu")
        assert int(util.ulp_diff(c, cc).max()) <= 1
        d = O.pyr_down_gauss_f32(d, impl="cpu")


def _icp_inputs(pair, holes=True):
    K, pose0, pose1, f0, f1 = pair
    depth = util.punch_holes(f1["depth"]) if holes else f1["depth"]
    v4, n4 = util.holes_in_maps(f0["vmap"], f0["nmap"]) if holes else (f0["vmap"], f0["nmap"])
    R = pose0[:3, :3].astype(np.float32)
    t = pose0[:3, 3].astype(np.float32)
    vp, npv = O.copy_maps(v4, n4, impl="cpu")
    levels = []
    for level in range(3):
        fx, fy, cx, cy = util.se3_level_params(K, level)
        vc = O.create_vmap(depth, fx, fy, cx, cy, 20.0, impl="cpu")
        nc = O.create_nmap(vc, impl="cpu")
        gvp, gnp = O.transform_maps(vp, npv, R, t, impl="cpu")
        levels.append((fx, fy, cx, cy, vc, nc, gvp, gnp))
        depth = O.pyr_down_u16(depth, impl="cpu")
        vp, npv = O.resize_map(vp, False, impl="cpu"), O.resize_map(npv, True, impl="cpu")
    return R, t, levels


def test_icp_step(ops, pair, have_ref):
    R, t, levels = _icp_inputs(pair)
    Rinv = np.linalg.inv(R.astype(np.float64)).astype(np.float32)
    ang = float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0)))
    for fx, fy, cx, cy, vc, nc, gvp, gnp in levels:
        args = (R, t, vc, nc, Rinv, t, fx, fy, cx, cy, gvp, gnp, 0.10, ang)
        A, b, res = ops.icpStep(*args)
        Ar, br, resr = O.icp_step(*args, impl="ref")
        Ac, bc, resc = O.icp_step(*args, impl="cpu")
        assert res[1] > 1000
        # inlier count is an exact small-integer float sum: must agree exactly with the reference kernels
        assert res[1] == resr[1]
        assert abs(res[1] - resc[1]) <= max(3.0, 2e-4 * resc[1])  # IEEE vs approx division may flip a pixel at a threshold
        for got, want in ((A, Ar), (b, br)):
            assert rel_err(got, want) < 1e-4
        # the CPU oracle accumulates in double: it is the better ground truth (tighter than the reference's own noise)
        assert rel_err(A, Ac) < 1e-4 and rel_err(b, bc) < 1e-3
        assert abs(res[0] - resc[0]) <= 1e-3 * resc[0]
        # determinism: same launch, same bits
        A2, b2, res2 = ops.icpStep(*args)
        assert np.array_equal(A, A2) and np.array_equal(b, b2) and np.array_equal(res, res2)


def _rgb_inputs(pair):
    K, pose0, pose1, f0, f1 = pair
    out = []
    ld = O.vertices_to_depth(f0["vmap"], 6.0, impl="cpu")
    nd = ld.copy()  # the frame-to-model flow builds nextDepth from the MODEL vertices (RGBDOdometry.cpp:245 quirk)
    li = O.bgr_to_intensity(f0["rgba"], impl="cpu")
    ni = O.bgr_to_intensity(f1["rgba"], impl="cpu")
    rel = np.linalg.inv(pose1) @ pose0  # maps last-frame points into the next frame; the kernels take its inverse
    Rt = np.linalg.inv(rel)
    for level in range(3):
        fx, fy, cx, cy = util.se3_level_params(K, level)
        Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        krk = (Km @ Rt[:3, :3] @ np.linalg.inv(Km)).astype(np.float32)
        kt = (Km @ Rt[:3, 3]).astype(np.float32)
        dx, dy = O.derivative_images(ni, impl="cpu")
        out.append((level, fx, fy, cx, cy, krk, kt, dx, dy, ld, nd, li, ni))
        ld, nd = O.pyr_down_gauss_f32(ld, impl="cpu"), O.pyr_down_gauss_f32(nd, impl="cpu")
        li, ni = O.pyr_down_gauss_u8(li, impl="cpu"), O.pyr_down_gauss_u8(ni, impl="cpu")
    return K, out


def test_rgb_residual_and_step(ops, pair, have_ref):
    K, levels = _rgb_inputs(pair)
    min_grad = [5, 3, 1]
    for level, fx, fy, cx, cy, krk, kt, dx, dy, ld, nd, li, ni in levels:
        min_scale = float(min_grad[level] ** 2 / 0.125 ** 2)
        cor, sig, cnt = ops.computeRgbResidual(min_scale, dx, dy, ld, nd, li, ni, 0.07, kt, krk)
        corr, sigr, cntr = O.rgb_residual(min_scale, dx, dy, ld, nd, li, ni, 0.07, kt, krk, impl="ref")
        corc, sigc, cntc = O.rgb_residual(min_scale, dx, dy, ld, nd, li, ni, 0.07, kt, krk, impl="cpu")
        assert cnt > 500, (level, cnt)
        assert (cnt, sig) == (cntr, sigr)
        assert np.array_equal(cor["valid"] != 0, corr["valid"] != 0)
        m = cor["valid"] != 0
        for f in ("zero_x", "zero_y", "one_x", "one_y", "diff"):
            assert np.array_equal(cor[f][m], corr[f][m]), f
        assert abs(cnt - cntc) <= max(2, 1e-3 * cntc)

        cloud = O.project_point_cloud(ld, K.fx, K.fy, K.cx, K.cy, level, impl="cpu")
        for sigma in (float(np.sqrt(cnt)), -1.0):
            A, b = ops.rgbStep(corr, sigma, cloud, fx, fy, dx, dy, 0.125)
            Ar, br = O.rgb_step(corr, sigma, cloud, fx, fy, dx, dy, 0.125, impl="ref")
            Ac, bc = O.rgb_step(corr, sigma, cloud, fx, fy, dx, dy, 0.125, impl="cpu")
            assert rel_err(A, Ar) < 1e-4 and rel_err(b, br) < 1e-4
            assert rel_err(A, Ac) < 1e-4 and rel_err(b, bc) < 1e-3


def test_so3_step(ops, pair, have_ref):
    K, levels = _rgb_inputs(pair)
    level, fx, fy, cx, cy, _, _, _, _, _, _, li, ni = levels[2]
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    for ang in (0.0, 0.004):
        c, s = np.cos(ang), np.sin(ang)
        Rr = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        H = (Km @ Rr @ np.linalg.inv(Km)).astype(np.float32)
        kinv = np.linalg.inv(Km).astype(np.float32)
        krlr = (Km @ Rr).astype(np.float32)
        A, b, res = ops.so3Step(li, ni, H, kinv, krlr)
        Ar, br, resr = O.so3_step(li, ni, H, kinv, krlr, impl="ref")
        Ac, bc, resc = O.so3_step(li, ni, H, kinv, krlr, impl="cpu")
        assert res[1] == resr[1] and res[1] > 1000
        assert rel_err(A, Ar) < 1e-4 and rel_err(b, br) < 1e-4 and abs(res[0] - resr[0]) <= 1e-4 * resr[0]
        assert rel_err(A, Ac) < 1e-4 and rel_err(b, bc) < 1e-3


def test_odd_sizes_scalar_path(ops, have_ref):
    """widths that are not a multiple of 4 take the scalar path; tiny images; fully invalid images."""
    rng = np.random.default_rng(5)
    depth = (rng.random((30, 42)) * 3000 + 500).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.1] = 0
    assert np.array_equal(ops.pyrDown(depth), O.pyr_down_u16(depth, impl="ref"))
    v = ops.createVMap(depth, 50.0, 50.0, 21.0, 15.0, 20.0)
    vr = O.create_vmap(depth, 50.0, 50.0, 21.0, 15.0, 20.0, impl="ref")
    assert util.masked_map_compare(v, vr, 30) == 0
    n = ops.createNMap(v)
    nr = O.create_nmap(vr, impl="ref")
    assert util.masked_map_compare(n, nr, 30) <= 1
    I = np.eye(3, dtype=np.float32)
    z = np.zeros(3, np.float32)
    A, b, res = ops.icpStep(I, z, v, n, I, z, 50.0, 50.0, 21.0, 15.0, v, n, 0.1, 0.342)
    Ar, br, resr = O.icp_step(I, z, vr, nr, I, z, 50.0, 50.0, 21.0, 15.0, vr, nr, 0.1, 0.342, impl="ref")
    assert res[1] == resr[1] and rel_err(A, Ar) < 1e-4
    # all-invalid input: zero system, zero inliers
    dz = np.zeros((32, 48), np.uint16)
    vz = ops.createVMap(dz, 50.0, 50.0, 24.0, 16.0, 20.0)
    assert np.isnan(vz[:32]).all()
    nz = ops.createNMap(vz)
    A, b, res = ops.icpStep(I, z, vz, nz, I, z, 50.0, 50.0, 24.0, 16.0, vz, nz, 0.1, 0.342)
    assert not A.any() and not b.any() and res[1] == 0
