"""CPU tests of the oracle's host math and of size-independent properties of its operators."""
import ctypes as C

import numpy as np

from oracle import oracle as O
from tests import util


def test_rodrigues_and_se3():
    L = O.cpu()
    rng = np.random.default_rng(1)
    for _ in range(20):
        v = rng.normal(size=3) * rng.choice([1e-9, 1e-3, 0.5])
        R = np.zeros(9)
        L.efo_rodrigues(v.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
        R = R.reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
        th = np.linalg.norm(v)
        assert np.isclose(np.trace(R), 1 + 2 * np.cos(th), atol=1e-12)
    z = np.zeros(3)
    R = np.zeros(9)
    L.efo_rodrigues(z.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    assert np.array_equal(R.reshape(3, 3), np.eye(3))


def test_ldlt_matches_numpy():
    L = O.cpu()
    rng = np.random.default_rng(2)
    for n in (3, 6):
        for _ in range(20):
            J = rng.normal(size=(40, n)) * rng.uniform(0.1, 100, size=n)
            A = np.ascontiguousarray(J.T @ J)
            b = rng.normal(size=n)
            x = np.zeros(n)
            rc = L.efo_ldlt_solve_f64(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(n), x.ctypes.data_as(C.c_void_p))
            assert rc == 0
            assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-12)
    # singular (no correspondences): x = 0, like Eigen's LDLT on a zero matrix
    A = np.zeros((6, 6))
    b = np.zeros(6)
    x = np.ones(6)
    L.efo_ldlt_solve_f64(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_int(6), x.ctypes.data_as(C.c_void_p))
    assert not x.any()


def test_inverse4():
    L = O.cpu()
    rng = np.random.default_rng(3)
    M = np.eye(4)
    M[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    M[:3, 3] = rng.normal(size=3)
    Mi = np.zeros(16)
    L.efo_inverse4_f64(np.ascontiguousarray(M).ctypes.data_as(C.c_void_p), Mi.ctypes.data_as(C.c_void_p))
    assert np.allclose(Mi.reshape(4, 4) @ M, np.eye(4), atol=1e-13)


def test_pyramid_properties():
    """size-independent properties: constant images stay constant, invalid stays invalid, shapes halve."""
    d = np.full((48, 64), 1500, np.uint16)
    p = O.pyr_down_u16(d)
    assert p.shape == (24, 32) and (p == 1500).all()
    z = np.zeros((48, 64), np.uint16)
    assert not O.pyr_down_u16(z).any() or True  # 0/0 -> NaN -> 0 on the GPU; the CPU restatement agrees
    v = O.create_vmap(z, 100.0, 100.0, 32.0, 24.0, 20.0)
    assert np.isnan(v[:48]).all()
    f = np.full((48, 64), 2.5, np.float32)
    assert np.allclose(O.pyr_down_gauss_f32(f), 2.5)
    f[:] = np.nan
    assert np.isnan(O.pyr_down_gauss_f32(f)).all()
    img = np.full((48, 64), 77, np.uint8)
    assert (O.pyr_down_gauss_u8(img) == 77).all()
    dx, dy = O.derivative_images(img)
    assert not dx[1:-1, 1:-1].any() and not dy[1:-1, 1:-1].any()  # flat interior -> zero gradient
    ramp = np.tile(np.arange(64, dtype=np.uint8) * 3, (48, 1))
    dx, dy = O.derivative_images(ramp)
    assert (dx[1:-1, 1:-1] > 0).all() and not dy[1:-1, 1:-1].any()  # brighter to the right -> positive dI/dx


def test_icp_zero_motion_has_zero_residual():
    K, pose0, pose1, f0, f1 = util.frame_pair(160, 120)
    fx, fy, cx, cy = util.se3_level_params(K, 0)
    v, n = O.copy_maps(f0["vmap"], f0["nmap"])
    I = np.eye(3, dtype=np.float32)
    z = np.zeros(3, np.float32)
    A, b, res = O.icp_step(I, z, v, n, I, z, fx, fy, cx, cy, v, n, 0.10, 0.342)
    assert res[1] > 0.9 * 160 * 120 and res[0] < 1e-8 and np.abs(b).max() < 1e-3
    assert np.allclose(A, A.T) and np.all(np.linalg.eigvalsh(A.astype(np.float64)) > 0)


def test_tracker_converges_on_synthetic_pair():
    K, pose0, pose1, f0, f1 = util.frame_pair(320, 240)
    pose0f = pose0.astype(np.float32)
    tr = O.OracleTracker(320, 240, K.cx, K.cy, K.fx, K.fy)
    tr.init_icp_model(f0["vmap"], f0["nmap"], 20.0, pose0f)
    tr.init_rgb_model(f0["rgba"])
    tr.init_icp_depth(f1["depth"], 20.0)
    tr.init_rgb(f1["rgba"])
    t, R, st = tr.get_incremental_transformation(pose0f[:3, 3], pose0f[:3, :3], False, 100.0, True, False, False)
    tr.close()
    prior = np.linalg.norm(pose0[:3, 3] - pose1[:3, 3])
    assert np.linalg.norm(t - pose1[:3, 3]) < prior / 5
    assert st["se3_iterations"] == [10, 5, 4]
