"""ctypes binding of libef_track.so (include/ef_track.h).  Fails loudly if the library is missing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# EF_TRACK_LIB: developer override (kernel-variant sweeps, tools/build_variant.sh); the product is libef_track.so
_SO = os.environ.get("EF_TRACK_LIB") or os.path.join(_HERE, "libef_track.so")
_lib = None


class EFError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where} failed with code {code}" + (f": {detail}" if detail else ""))


class TrackStats(C.Structure):
    """ef_track_stats (include/ef_track.h)"""
    _fields_ = [("last_icp_error", C.c_float), ("last_icp_count", C.c_float),
                ("last_rgb_error", C.c_float), ("last_rgb_count", C.c_float),
                ("last_so3_error", C.c_float), ("last_so3_count", C.c_float),
                ("last_A", C.c_double * 36), ("last_b", C.c_double * 6),
                ("so3_iterations", C.c_int), ("se3_iterations", C.c_int * 3)]


class StageTimes(C.Structure):
    """ef_stage_times (include/ef_track.h): the reference's Stopwatch keys"""
    _fields_ = [(n, C.c_float) for n in ("so3_step_ms", "rgb_residual_ms", "icp_step_ms", "rgb_step_ms", "so3_step_sum_ms", "rgb_residual_sum_ms",
                                         "icp_step_sum_ms", "rgb_step_sum_ms", "iteration_ms", "iteration_sum_ms", "call_ms")] + [("solve_mode", C.c_int)]


class FrameInputs(C.Structure):
    """ef_frame_inputs (include/ef_track.h)"""
    _fields_ = [("vertices_rgba32f", C.c_void_p), ("normals_rgba32f", C.c_void_p), ("model_rgba8", C.c_void_p), ("depth", C.c_void_p),
                ("rgba8", C.c_void_p), ("depth_cutoff", C.c_float), ("on_host", C.c_int)]


def lib_path() -> str:
    return _SO


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=out)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise EFError(-2, "loading libef_track.so",
                          f"{_SO} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
        L = C.CDLL(_SO)
        L.ef_last_error.restype = C.c_char_p
        L.ef_tracker_stream.restype = C.c_void_p
        L.ef_tracker_launch_count.restype = C.c_longlong
        L.ef_op_scratch_bytes.restype = C.c_size_t
        L.ef_op_splat_scratch_bytes.restype = C.c_size_t
        L.ef_default_dist_thresh.restype = C.c_float
        L.ef_default_angle_thresh.restype = C.c_float
        _lib = L
    return _lib


# every symbol include/ef_track.h declares (checked by tests/test_abi.py against the header text)
EXPORTED = [
    "ef_tracker_create", "ef_tracker_destroy", "ef_tracker_set_option", "ef_tracker_get_option", "ef_last_error",
    "ef_tracker_stream", "ef_tracker_synchronize", "ef_tracker_wait_event", "ef_tracker_wait_stream", "ef_default_dist_thresh", "ef_default_angle_thresh",
    "ef_init_icp_depth", "ef_init_icp_depth_raw", "ef_init_icp_depth_raw_host", "ef_init_icp_maps", "ef_init_icp_model", "ef_init_rgb", "ef_init_rgb_model", "ef_init_first_rgb",
    "ef_init_icp_depth_array", "ef_init_icp_maps_array", "ef_init_icp_model_array", "ef_init_rgb_array",
    "ef_init_rgb_model_array", "ef_init_first_rgb_array",
    "ef_init_icp_depth_host", "ef_init_icp_maps_host", "ef_init_icp_model_host", "ef_init_rgb_host", "ef_init_rgb_model_host",
    "ef_init_first_rgb_host",
    "ef_get_incremental_transformation", "ef_get_incremental_transformation_launch",
    "ef_get_incremental_transformation_finish", "ef_track_frame_to_model_launch", "ef_track_frame_to_model", "ef_batch_width", "ef_track_frames_to_model_batch_launch",
    "ef_track_frames_to_model_batch", "ef_get_covariance", "ef_tracker_download", "ef_tracker_launch_count", "ef_tracker_profile", "ef_tracker_trace", "ef_tracker_stage_times",
    "ef_op_pyr_down_u16", "ef_op_create_vmap", "ef_op_create_nmap", "ef_op_transform_maps", "ef_op_copy_maps",
    "ef_op_resize_map", "ef_op_vertices_to_depth", "ef_op_pyr_down_gauss_f32", "ef_op_pyr_down_gauss_u8",
    "ef_op_bgr_to_intensity", "ef_op_depth_bilateral", "ef_op_depth_metric", "ef_op_derivative_images", "ef_op_project_point_cloud", "ef_op_icp_step",
    "ef_op_rgb_residual", "ef_op_rgb_step", "ef_op_so3_step", "ef_op_scratch_bytes", "ef_abi_version", "ef_device_count",
    "ef_op_splat_scratch_bytes", "ef_op_splat_predict", "ef_op_splat_predict_inst", "ef_op_fill_vertex", "ef_op_fill_normal", "ef_op_fill_rgb",
]
