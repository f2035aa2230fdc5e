"""Generate tests/golden/ref_cuda_160x120.npz: outputs of the REFERENCE's own CUDA kernels
(oracle/_ref/libef_ref.so = elasticfusionpublic/Core/src/Cuda/{cudafuncs.cu,reduce.cu} compiled unmodified,
see oracle/Makefile) on seeded synthetic inputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_cuda_160x120.npz'

then copy the file into tests/golden/.  The inputs are stored next to the outputs so that the CPU-side test
(tests/test_oracle_golden.py) does not depend on bit-reproducible trigonometry in the scene generator.
The reference publishes no golden vectors of its own for this path; these pin the C
restatement oracle/ef_oracle.c to what the reference code really computes on a B200.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import util  # noqa: E402

W, H = 160, 120


def main(out_path):
    K, pose0, pose1, f0, f1 = util.frame_pair(W, H)
    depth = util.punch_holes(f1["depth"], frac=0.03)
    v4, n4 = util.holes_in_maps(f0["vmap"], f0["nmap"])
    rgba0, rgba1 = f0["rgba"].copy(), f1["rgba"].copy()
    rgba1[40:46, 60:80] = 0
    g = {"K": np.array([K.fx, K.fy, K.cx, K.cy], np.float32), "pose0": pose0, "pose1": pose1, "depth1": depth, "vmap0": v4, "nmap0": n4,
         "rgba0": rgba0, "rgba1": rgba1}
    fx, fy, cx, cy = [np.float32(x) for x in (K.fx, K.fy, K.cx, K.cy)]
    R = pose0[:3, :3].astype(np.float32)
    t = pose0[:3, 3].astype(np.float32)

    # ---- image / pyramid operators ----
    g["pyr_down_u16"] = O.pyr_down_u16(depth, impl="ref")
    g["vmap_curr"] = O.create_vmap(depth, fx, fy, cx, cy, 20.0, impl="ref")
    g["nmap_curr"] = O.create_nmap(g["vmap_curr"], impl="ref")
    vp, npv = O.copy_maps(v4, n4, impl="ref")
    g["copy_v"], g["copy_n"] = vp, npv
    g["resize_v"] = O.resize_map(vp, False, impl="ref")
    g["resize_n"] = O.resize_map(npv, True, impl="ref")
    g["transform_v"], g["transform_n"] = O.transform_maps(vp, npv, R, t, impl="ref")
    g["depth_f32"] = O.vertices_to_depth(v4, 6.0, impl="ref")
    g["pyr_down_gauss_f32"] = O.pyr_down_gauss_f32(g["depth_f32"], impl="ref")
    g["intensity0"] = O.bgr_to_intensity(rgba0, impl="ref")
    g["intensity1"] = O.bgr_to_intensity(rgba1, impl="ref")
    g["pyr_down_gauss_u8"] = O.pyr_down_gauss_u8(g["intensity1"], impl="ref")
    g["dIdx"], g["dIdy"] = O.derivative_images(g["intensity1"], impl="ref")
    g["cloud"] = O.project_point_cloud(g["depth_f32"], K.fx, K.fy, K.cx, K.cy, 0, impl="ref")

    # ---- association + reduction operators ----
    Rinv = np.linalg.inv(R.astype(np.float64)).astype(np.float32)
    ang = float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0)))
    A, b, res = O.icp_step(R, t, g["vmap_curr"], g["nmap_curr"], Rinv, t, fx, fy, cx, cy, g["transform_v"], g["transform_n"], 0.10, ang,
                           impl="ref")
    g["icp_A"], g["icp_b"], g["icp_res"] = A, b, res
    rel = np.linalg.inv(np.linalg.inv(pose1) @ pose0)
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    krk = (Km @ rel[:3, :3] @ np.linalg.inv(Km)).astype(np.float32)
    kt = (Km @ rel[:3, 3]).astype(np.float32)
    g["krkinv"], g["kt"] = krk, kt
    cor, sig, cnt = O.rgb_residual(64.0, g["dIdx"], g["dIdy"], g["depth_f32"], g["depth_f32"], g["intensity0"], g["intensity1"], 0.07, kt,
                                   krk, impl="ref")
    g["rgbres_valid"] = (cor["valid"] != 0)
    g["rgbres_zero_x"], g["rgbres_zero_y"], g["rgbres_diff"] = cor["zero_x"], cor["zero_y"], cor["diff"]
    g["rgbres_sigma_count"] = np.array([sig, cnt], np.int64)
    for name, sigma in (("w", float(np.sqrt(max(cnt, 1)))), ("unit", -1.0)):
        A, b = O.rgb_step(cor, sigma, g["cloud"], fx, fy, g["dIdx"], g["dIdy"], 0.125, impl="ref")
        g["rgb_A_" + name], g["rgb_b_" + name] = A, b
    g["rgb_sigma_w"] = np.float32(np.sqrt(max(cnt, 1)))
    c, s = np.cos(0.004), np.sin(0.004)
    Rr = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    g["so3_H"] = (Km @ Rr @ np.linalg.inv(Km)).astype(np.float32)
    g["so3_kinv"] = np.linalg.inv(Km).astype(np.float32)
    g["so3_krlr"] = (Km @ Rr).astype(np.float32)
    A, b, res = O.so3_step(g["intensity0"], g["intensity1"], g["so3_H"], g["so3_kinv"], g["so3_krlr"], impl="ref")
    g["so3_A"], g["so3_b"], g["so3_res"] = A, b, res

    # ---- full tracker (reference operators under the restated host loop) ----
    pose0f = pose0.astype(np.float32)
    modes = {"icp_nopyr": (False, 100.0, False, False, False), "joint": (False, 10.0, True, False, False), "joint_so3": (False, 10.0, True, False, True)}
    tr = O.OracleTracker(W, H, K.cx, K.cy, K.fx, K.fy, impl="ref")
    for name, (rgb_only, w, pyr, fast, so3) in modes.items():
        tr.init_first_rgb(rgba0)
        tr.init_icp_model(v4, n4, 20.0, pose0f)
        tr.init_rgb_model(rgba0)
        tr.init_icp_depth(depth, 20.0)
        tr.init_rgb(rgba1)
        tt, RR, st = tr.get_incremental_transformation(pose0f[:3, 3], pose0f[:3, :3], rgb_only, w, pyr, fast, so3)
        g["track_%s_t" % name], g["track_%s_R" % name] = tt, RR
        g["track_%s_stats" % name] = np.array([st["last_icp_error"], st["last_icp_count"], st["last_rgb_error"], st["last_rgb_count"],
                                               st["last_so3_error"], st["last_so3_count"]], np.float64)
        g["track_%s_iters" % name] = np.array(st["se3_iterations"] + [st["so3_iterations"]], np.int64)
        g["track_%s_A" % name] = st["last_A"]
    tr.close()
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, os.path.getsize(out_path), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_cuda_160x120.npz"))
