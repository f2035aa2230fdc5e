"""Python mirror of class RGBDOdometry (elasticfusionpublic/Core/src/Utils/RGBDOdometry.h:31-134)
over the C ABI.  Same method names, argument meaning and call-order contract as the reference; a
GPUTexture* argument becomes either a CUDA torch tensor (device pointer borrowed for the call) or a
host numpy array / CPU tensor (copied through the handle's staging buffers, `_host` entry points).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding
from .binding import EFError, TrackStats

EF_OPT_SOLVE_MODE, EF_OPT_USE_GRAPH, EF_OPT_FUSED_BUILD, EF_OPT_PROFILE, EF_OPT_GRID_CTAS, EF_OPT_AUX_STREAMS, EF_OPT_FRAME_BUILD, EF_OPT_DEFER_BUILD, \
    EF_OPT_HOST_FUSED = 1, 2, 3, 4, 5, 6, 7, 8, 9
EF_SOLVE_HOST, EF_SOLVE_DEVICE = 0, 1

_LEVEL_BUFFERS = {"vmap_curr": (np.float32, 3), "nmap_curr": (np.float32, 3), "vmap_g_prev": (np.float32, 3),
                  "nmap_g_prev": (np.float32, 3), "last_depth": (np.float32, 1), "next_depth": (np.float32, 1),
                  "last_image": (np.uint8, 1), "next_image": (np.uint8, 1), "last_next_image": (np.uint8, 1),
                  "dIdx": (np.int16, 1), "dIdy": (np.int16, 1), "depth_tmp": (np.uint16, 1), "filt_depth": (np.uint16, 1)}


def _is_cuda_tensor(a) -> bool:
    return hasattr(a, "is_cuda") and a.is_cuda


def _is_cuda_array(a) -> bool:
    """a cudaArray_t wrapper (attribute `cuda_array_handle`): the GL-interop form of a GPUTexture (RGBDOdometry.cpp:120-128)"""
    return hasattr(a, "cuda_array_handle")


def _arr(a):
    return C.c_void_p(a.cuda_array_handle)


def _dev_ptr(a, dtype_name: str):
    import torch
    want = getattr(torch, dtype_name)
    if a.dtype != want or not a.is_contiguous():
        raise ValueError(f"expected a contiguous {dtype_name} CUDA tensor, got {a.dtype}")
    return C.c_void_p(a.data_ptr())


_raw_stream = None


def _torch_raw_stream():
    """torch's current CUDA stream as a raw cudaStream_t (an int), without building a torch.cuda.Stream object per call"""
    global _raw_stream
    import torch
    if _raw_stream is None:
        fast = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        _raw_stream = (lambda: fast(torch.cuda.current_device())) if fast else (lambda: torch.cuda.current_stream().cuda_stream)
    return _raw_stream()


def _host_arr(a, dtype):
    if hasattr(a, "numpy"):
        a = a.numpy()
    return np.ascontiguousarray(a, dtype=dtype)


class RGBDOdometry:
    """RGBDOdometry(width, height, cx, cy, fx, fy, distThresh=0.10, angleThresh=sin(20 deg))"""

    def __init__(self, width, height, cx, cy, fx, fy, distThresh=None, angleThresh=None, stream=None,
                 solve_mode=None):
        self._L = binding.lib()
        if distThresh is None:
            distThresh = self._L.ef_default_dist_thresh()
        if angleThresh is None:
            angleThresh = self._L.ef_default_angle_thresh()
        self.width, self.height = int(width), int(height)
        h = C.c_void_p()
        rc = self._L.ef_tracker_create(C.c_int(width), C.c_int(height), C.c_float(cx), C.c_float(cy), C.c_float(fx),
                                       C.c_float(fy), C.c_float(distThresh), C.c_float(angleThresh),
                                       C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise EFError(rc, "ef_tracker_create", "no CUDA device (no CPU fallback)" if rc == -2 else "")
        self._h = h
        self._stats = TrackStats()
        # initial values of the reference's public fields (RGBDOdometry.cpp:26-31)
        self._stats.last_icp_count = self._stats.last_rgb_count = self._stats.last_so3_count = float(width * height)
        # result buffers of finish() and their addresses, made once (ctypes conversions are the bulk of a call's host time)
        self._res = np.zeros(12, np.float32)
        self._res_t, self._res_r = C.c_void_p(self._res.ctypes.data), C.c_void_p(self._res.ctypes.data + 12)
        self._stats_ref = C.byref(self._stats)
        # argument buffers of the per-frame calls, with their addresses (numpy's .ctypes costs ~2 us per access)
        self._pose = np.zeros(16, np.float32)
        self._pose_p = C.c_void_p(self._pose.ctypes.data)
        self._prior = np.zeros(12, np.float32)
        self._prior_t, self._prior_r = C.c_void_p(self._prior.ctypes.data), C.c_void_p(self._prior.ctypes.data + 12)
        self._frame = None
        self._stream_int = int(self._L.ef_tracker_stream(self._h) or 0)
        if solve_mode is not None:  # None: the library's default (EF_SOLVE_DEVICE wherever the image fits the persistent kernel)
            self.set_option(EF_OPT_SOLVE_MODE, solve_mode)

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, rc, where):
        if rc != 0:
            raise EFError(rc, where, self._L.ef_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.ef_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self._L.ef_tracker_set_option(self._h, C.c_int(key), C.c_int(value)), "ef_tracker_set_option")

    def get_option(self, key):
        v = C.c_int(0)
        self._check(self._L.ef_tracker_get_option(self._h, C.c_int(key), C.byref(v)), "ef_tracker_get_option")
        return v.value

    def _borrow(self):
        """Stream contract of include/ef_track.h: device inputs are consumed on the HANDLE's stream.  If torch's current stream
        still has work in flight (it may be producing the tensor we are about to borrow), the handle's stream waits for it
        (ef_tracker_wait_stream: one C call, nothing enqueued when that stream is idle)."""
        rc = self._L.ef_tracker_wait_stream(self._h, C.c_void_p(_torch_raw_stream()))
        if rc != 0:
            self._check(rc, "ef_tracker_wait_stream")

    @property
    def stream(self):
        return self._L.ef_tracker_stream(self._h)

    @property
    def launch_count(self):
        return int(self._L.ef_tracker_launch_count(self._h))

    def profile(self):
        """(solve_ms_total, calls) accumulated since the last read (needs set_option(EF_OPT_PROFILE, 1))."""
        ms, n = C.c_double(0), C.c_longlong(0)
        self._check(self._L.ef_tracker_profile(self._h, C.byref(ms), C.byref(n)), "ef_tracker_profile")
        return ms.value, n.value

    def stageTimes(self):
        """{"so3Step": ms, "computeRgbResidual": ms, "icpStep": ms, "rgbStep": ms, ...}: the reference's Stopwatch keys
        (Utils/RGBDOdometry.cpp:333-538) for the last getIncrementalTransformation (see ef_stage_times in include/ef_track.h)"""
        st = binding.StageTimes()
        self._check(self._L.ef_tracker_stage_times(self._h, C.byref(st)), "ef_tracker_stage_times")
        return {"so3Step": st.so3_step_ms, "computeRgbResidual": st.rgb_residual_ms, "icpStep": st.icp_step_ms, "rgbStep": st.rgb_step_ms,
                "so3Step_sum": st.so3_step_sum_ms, "computeRgbResidual_sum": st.rgb_residual_sum_ms, "icpStep_sum": st.icp_step_sum_ms,
                "rgbStep_sum": st.rgb_step_sum_ms, "iteration": st.iteration_ms, "iteration_sum": st.iteration_sum_ms, "call": st.call_ms,
                "solve_mode": st.solve_mode}

    def synchronize(self):
        self._check(self._L.ef_tracker_synchronize(self._h), "ef_tracker_synchronize")

    # -- the reference API ------------------------------------------------------------------------
    def initICP(self, *args):
        """initICP(filteredDepth, depthCutoff)  or  initICP(predictedVertices, predictedNormals, depthCutoff)"""
        if len(args) == 2:
            depth, cutoff = args
            if _is_cuda_array(depth):
                self._check(self._L.ef_init_icp_depth_array(self._h, _arr(depth), C.c_float(cutoff)), "ef_init_icp_depth_array")
            elif _is_cuda_tensor(depth):
                self._borrow()
                self._check(self._L.ef_init_icp_depth(self._h, _dev_ptr(depth, "uint16"), C.c_size_t(0), C.c_float(cutoff)),
                            "ef_init_icp_depth")
            else:
                d = _host_arr(depth, np.uint16)
                self._keep = d
                self._check(self._L.ef_init_icp_depth_host(self._h, d.ctypes.data_as(C.c_void_p), C.c_float(cutoff)),
                            "ef_init_icp_depth_host")
        elif len(args) == 3:
            v, n, cutoff = args
            if _is_cuda_array(v):
                self._check(self._L.ef_init_icp_maps_array(self._h, _arr(v), _arr(n), C.c_float(cutoff)), "ef_init_icp_maps_array")
            elif _is_cuda_tensor(v):
                self._borrow()
                self._check(self._L.ef_init_icp_maps(self._h, _dev_ptr(v, "float32"), _dev_ptr(n, "float32"), C.c_float(cutoff)),
                            "ef_init_icp_maps")
            else:
                hv, hn = _host_arr(v, np.float32), _host_arr(n, np.float32)
                self._keep = (hv, hn)
                self._check(self._L.ef_init_icp_maps_host(self._h, hv.ctypes.data_as(C.c_void_p), hn.ctypes.data_as(C.c_void_p),
                                                          C.c_float(cutoff)), "ef_init_icp_maps_host")
        else:
            raise TypeError("initICP takes (depth, cutoff) or (vertices, normals, cutoff)")

    def initICPRaw(self, rawDepth, maxDepth, depthCutoff):
        """filterDepth + initICP (ElasticFusion.cpp:309, :348): RAW sensor depth (u16 mm) -> bilateral filter -> pyramids."""
        if _is_cuda_tensor(rawDepth):
            self._borrow()
            self._check(self._L.ef_init_icp_depth_raw(self._h, _dev_ptr(rawDepth, "uint16"), C.c_size_t(0), C.c_float(maxDepth),
                                                      C.c_float(depthCutoff)), "ef_init_icp_depth_raw")
        else:
            d = _host_arr(rawDepth, np.uint16)
            self._keep = d
            self._check(self._L.ef_init_icp_depth_raw_host(self._h, d.ctypes.data_as(C.c_void_p), C.c_float(maxDepth),
                                                           C.c_float(depthCutoff)), "ef_init_icp_depth_raw_host")

    def initICPModel(self, predictedVertices, predictedNormals, depthCutoff, modelPose):
        self._pose[:] = np.asarray(modelPose).reshape(16)
        if _is_cuda_array(predictedVertices):
            self._check(self._L.ef_init_icp_model_array(self._h, _arr(predictedVertices), _arr(predictedNormals), C.c_float(depthCutoff),
                                                        self._pose_p), "ef_init_icp_model_array")
        elif _is_cuda_tensor(predictedVertices):
            self._borrow()
            self._check(self._L.ef_init_icp_model(self._h, _dev_ptr(predictedVertices, "float32"), _dev_ptr(predictedNormals, "float32"),
                                                  C.c_float(depthCutoff), self._pose_p), "ef_init_icp_model")
        else:
            hv, hn = _host_arr(predictedVertices, np.float32), _host_arr(predictedNormals, np.float32)
            self._keep = (hv, hn)
            self._check(self._L.ef_init_icp_model_host(self._h, hv.ctypes.data_as(C.c_void_p), hn.ctypes.data_as(C.c_void_p),
                                                       C.c_float(depthCutoff), self._pose_p), "ef_init_icp_model_host")

    def _rgb(self, name, rgb):
        if _is_cuda_array(rgb):
            self._check(getattr(self._L, name + "_array")(self._h, _arr(rgb)), name + "_array")
        elif _is_cuda_tensor(rgb):
            self._borrow()
            self._check(getattr(self._L, name)(self._h, _dev_ptr(rgb, "uint8"), C.c_size_t(0)), name)
        else:
            h = _host_arr(rgb, np.uint8)
            self._keep_rgb = h
            self._check(getattr(self._L, name + "_host")(self._h, h.ctypes.data_as(C.c_void_p)), name + "_host")

    def initRGB(self, rgb):
        self._rgb("ef_init_rgb", rgb)

    def initRGBModel(self, rgb):
        self._rgb("ef_init_rgb_model", rgb)

    def initFirstRGB(self, rgb):
        self._rgb("ef_init_first_rgb", rgb)

    # the public fields below read the stats block of the last call when they are asked for
    # public fields of the reference class (RGBDOdometry.h:64-73)
    lastICPError = property(lambda self: self._stats.last_icp_error)
    lastICPCount = property(lambda self: self._stats.last_icp_count)
    lastRGBError = property(lambda self: self._stats.last_rgb_error)
    lastRGBCount = property(lambda self: self._stats.last_rgb_count)
    lastSO3Error = property(lambda self: self._stats.last_so3_error)
    lastSO3Count = property(lambda self: self._stats.last_so3_count)
    so3_iterations = property(lambda self: self._stats.so3_iterations)
    se3_iterations = property(lambda self: list(self._stats.se3_iterations))

    # lastA / lastb of the last call (RGBDOdometry.h:72-73), unpacked from the stats block when asked for
    @property
    def lastA(self):
        return np.array(self._stats.last_A[:]).reshape(6, 6)

    @property
    def lastb(self):
        return np.array(self._stats.last_b[:])

    def getIncrementalTransformation(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        """trans (3,) and rot (3,3 row-major) are the pose prior; returns the updated (trans, rot)."""
        self._prior[:3] = np.asarray(trans).reshape(3)
        self._prior[3:] = np.asarray(rot).reshape(9)
        rc = self._L.ef_get_incremental_transformation(self._h, self._prior_t, self._prior_r, int(rgbOnly), C.c_float(icpWeight), int(pyramid),
                                                       int(fastOdom), int(so3), self._stats_ref)
        if rc != 0:
            self._check(rc, "ef_get_incremental_transformation")
        res = self._prior.copy()  # (in / out arguments of the C call)
        return res[:3], res[3:].reshape(3, 3)

    def launch(self, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        self._prior[:3] = np.asarray(trans).reshape(3)
        self._prior[3:] = np.asarray(rot).reshape(9)
        rc = self._L.ef_get_incremental_transformation_launch(self._h, self._prior_t, self._prior_r, int(rgbOnly), C.c_float(icpWeight),
                                                              int(pyramid), int(fastOdom), int(so3))
        if rc != 0:
            self._check(rc, "ef_get_incremental_transformation_launch")

    def finish(self):
        rc = self._L.ef_get_incremental_transformation_finish(self._h, self._res_t, self._res_r, self._stats_ref)
        if rc != 0:
            self._check(rc, "ef_get_incremental_transformation_finish")
        res = self._res.copy()
        return res[:3], res[3:].reshape(3, 3)

    # -- ElasticFusion::processFrame's frameToModel sequence (ElasticFusion.cpp:343-368) in one C call ---------
    def _frame_inputs(self, vertices, normals, model_rgba, depth, rgba, depthCutoff):
        from .binding import FrameInputs
        # 0: all device, 1: all host, 2: model maps on the device, sensor frame (depth, rgba) on the host
        model_dev, sensor_dev = _is_cuda_tensor(vertices), _is_cuda_tensor(depth)
        if model_dev and not (_is_cuda_tensor(normals) and _is_cuda_tensor(model_rgba)):
            raise ValueError("vertices, normals and the model image must live on the same side")
        if sensor_dev != _is_cuda_tensor(rgba) or (sensor_dev and not model_dev):
            raise ValueError("supported: all device, all host, or model maps on the device with the sensor frame on the host")
        on_host = 0 if sensor_dev else (2 if model_dev else 1)
        ptr = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        if model_dev:
            self._borrow()
        self._keep_frame = (vertices, normals, model_rgba, depth, rgba)
        fi = self._frame
        if fi is None:
            fi = self._frame = FrameInputs()
            self._frame_ref = C.byref(fi)
        fi.vertices_rgba32f, fi.normals_rgba32f, fi.model_rgba8 = ptr(vertices), ptr(normals), ptr(model_rgba)
        fi.depth, fi.rgba8, fi.depth_cutoff, fi.on_host = ptr(depth), ptr(rgba), float(depthCutoff), int(on_host)
        return fi

    def trackFrameToModelLaunch(self, vertices, normals, model_rgba, depth, rgba, depthCutoff, modelPose, rgbOnly, icpWeight, pyramid,
                                fastOdom, so3):
        self._frame_inputs(vertices, normals, model_rgba, depth, rgba, depthCutoff)
        self._pose[:] = np.asarray(modelPose).reshape(16)  # (raises unless it is a 4x4 matrix)
        rc = self._L.ef_track_frame_to_model_launch(self._h, self._frame_ref, self._pose_p, int(rgbOnly), C.c_float(icpWeight),
                                                    int(pyramid), int(fastOdom), int(so3))
        if rc != 0:
            self._check(rc, "ef_track_frame_to_model_launch")

    def trackFrameToModel(self, vertices, normals, model_rgba, depth, rgba, depthCutoff, modelPose, rgbOnly, icpWeight, pyramid, fastOdom,
                          so3):
        # one FFI call (ef_track_frame_to_model = launch + finish)
        self._frame_inputs(vertices, normals, model_rgba, depth, rgba, depthCutoff)
        self._pose[:] = np.asarray(modelPose).reshape(16)  # (raises unless it is a 4x4 matrix)
        rc = self._L.ef_track_frame_to_model(self._h, self._frame_ref, self._pose_p, self._res_t, self._res_r, int(rgbOnly),
                                             C.c_float(icpWeight), int(pyramid), int(fastOdom), int(so3), self._stats_ref)
        if rc != 0:
            self._check(rc, "ef_track_frame_to_model")
        res = self._res.copy()
        return res[:3], res[3:].reshape(3, 3)

    def getCovariance(self):
        cov = np.zeros(36)
        self._check(self._L.ef_get_covariance(self._h, cov.ctypes.data_as(C.c_void_p)), "ef_get_covariance")
        return cov.reshape(6, 6)

    # -- diagnostics ------------------------------------------------------------------------------
    def buffer(self, name, level):
        dt, planes = _LEVEL_BUFFERS[name]
        out = np.zeros(((self.height >> level) * planes, self.width >> level), dt)
        self._check(self._L.ef_tracker_download(self._h, name.encode(), C.c_int(level), out.ctypes.data_as(C.c_void_p),
                                                C.c_size_t(out.nbytes)), "ef_tracker_download")
        return out


class BatchTracker:
    """k independent sequences per kernel launch (ef_track_frames_to_model_batch): `trackers` = ef_batch_width() RGBDOdometry
    handles of one image size.  track(frames, poses, ...) takes one frameToModel frame per handle -- frames[g] =
    (vertices, normals, model_rgba, depth, rgba) as for RGBDOdometry.trackFrameToModel -- and returns [(trans, rot)] per handle;
    the handles' public fields (lastICPError, lastA, ...) are updated as by their own calls."""

    def __init__(self, trackers):
        self._L = binding.lib()
        self.n = len(trackers)
        if not 2 <= self.n <= self._L.ef_batch_width():
            raise ValueError(f"a batch holds 2 .. {self._L.ef_batch_width()} trackers")
        self.trackers = list(trackers)
        self._handles = (C.c_void_p * self.n)(*[t._h for t in self.trackers])
        self._inputs = (binding.FrameInputs * self.n)()
        self._poses = np.zeros((self.n, 16), np.float32)
        self._trans = np.zeros((self.n, 3), np.float32)
        self._rot = np.zeros((self.n, 9), np.float32)
        self._stats = (TrackStats * self.n)()
        self._poses_p = C.c_void_p(self._poses.ctypes.data)
        self._trans_p, self._rot_p = C.c_void_p(self._trans.ctypes.data), C.c_void_p(self._rot.ctypes.data)
        self._keep = None

    def _fill(self, frames, poses, depthCutoff):
        self._keep = frames
        for g, (tr, f) in enumerate(zip(self.trackers, frames)):
            fi = tr._frame_inputs(*f, depthCutoff)
            C.memmove(C.byref(self._inputs[g]), C.byref(fi), C.sizeof(binding.FrameInputs))
            self._poses[g] = np.asarray(poses[g], np.float32).reshape(16)

    def launch(self, frames, poses, depthCutoff, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        self._fill(frames, poses, depthCutoff)
        rc = self._L.ef_track_frames_to_model_batch_launch(self._handles, self.n, self._inputs, self._poses_p, int(rgbOnly),
                                                           C.c_float(icpWeight), int(pyramid), int(fastOdom), int(so3))
        if rc != 0:
            raise EFError(rc, "ef_track_frames_to_model_batch_launch", "; ".join(self._L.ef_last_error(t._h).decode() for t in self.trackers))

    def finish(self):
        return [t.finish() for t in self.trackers]

    def track(self, frames, poses, depthCutoff, rgbOnly, icpWeight, pyramid, fastOdom, so3):
        self._fill(frames, poses, depthCutoff)
        rc = self._L.ef_track_frames_to_model_batch(self._handles, self.n, self._inputs, self._poses_p,
                                                    self._trans_p, self._rot_p, int(rgbOnly),
                                                    C.c_float(icpWeight), int(pyramid), int(fastOdom), int(so3), self._stats)
        if rc != 0:
            raise EFError(rc, "ef_track_frames_to_model_batch", "; ".join(self._L.ef_last_error(t._h).decode() for t in self.trackers))
        for g, t in enumerate(self.trackers):
            C.memmove(C.byref(t._stats), C.byref(self._stats[g]), C.sizeof(TrackStats))
        return [(self._trans[g].copy(), self._rot[g].reshape(3, 3).copy()) for g in range(self.n)]
