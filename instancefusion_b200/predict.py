"""Device-resident mirror of the reference's model prediction (IndexMap::combinedPredict + FillIn, elasticfusionpublic/
Core/src/IndexMap.cpp:468-575, ElasticFusion.cpp:729-763) over ef_op_splat_predict / ef_op_fill_* of the C ABI: the surfel
buffer and the predicted maps stay in device memory (torch tensors, plumbing only), so the maps can go straight into
RGBDOdometry.trackFrameToModel without touching the host."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import binding
from .binding import EFError


class ModelPredictor:
    """predict(surfels, pose, time, ...) -> (image rgba8, vertex rgba32f, normal rgba32f, time u16) CUDA tensors, reused per call."""

    def __init__(self, width, height, cx, cy, fx, fy):
        self._L = binding.lib()
        self.width, self.height = int(width), int(height)
        self.cam = (float(cx), float(cy), float(fx), float(fy))
        dev = "cuda"
        self._keys = torch.empty(self._L.ef_op_splat_scratch_bytes(self.height, self.width), dtype=torch.uint8, device=dev)
        self.image = torch.empty((self.height, self.width, 4), dtype=torch.uint8, device=dev)
        self.vertex = torch.empty((self.height, self.width, 4), dtype=torch.float32, device=dev)
        self.normal = torch.empty((self.height, self.width, 4), dtype=torch.float32, device=dev)
        self.time = torch.empty((self.height, self.width), dtype=torch.uint16, device=dev)
        self.inst = torch.empty((self.height, self.width, 4), dtype=torch.uint8, device=dev)  # InstanceFusion's instance-colour target
        self.filled_vertex = torch.empty_like(self.vertex)
        self.filled_normal = torch.empty_like(self.normal)
        self.filled_image = torch.empty_like(self.image)

    def _stream(self):
        return C.c_void_p(self.stream if self.stream is not None else torch.cuda.current_stream().cuda_stream)

    stream = None  # a cudaStream_t (int) to run on, e.g. RGBDOdometry.stream; None = torch's current stream

    def predict(self, surfels: torch.Tensor, pose, time, maxTime=None, timeDelta=200, maxDepth=20.0, confThreshold=10.0):
        """IndexMap::combinedPredict(pose, model, depthCutoff, confThreshold, time, maxTime, timeDelta, ACTIVE)"""
        if not (surfels.is_cuda and surfels.dtype == torch.float32 and surfels.is_contiguous() and surfels.dim() == 2 and surfels.shape[1] >= 12):
            raise ValueError("surfels: contiguous float32 CUDA tensor of shape (N, >= 12)")
        t_inv = np.ascontiguousarray(np.linalg.inv(np.asarray(pose, np.float64)).astype(np.float32).reshape(16))
        cx, cy, fx, fy = self.cam
        rc = self._L.ef_op_splat_predict_inst(C.c_void_p(surfels.data_ptr()), C.c_size_t(surfels.shape[1] * 4), int(surfels.shape[0]),
                                         C.c_void_p(t_inv.ctypes.data), C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy), self.height,
                                         self.width, C.c_float(maxDepth), C.c_float(confThreshold), int(time),
                                         int(time if maxTime is None else maxTime), int(timeDelta), C.c_void_p(self._keys.data_ptr()),
                                         C.c_void_p(self.image.data_ptr()), C.c_void_p(self.vertex.data_ptr()),
                                         C.c_void_p(self.normal.data_ptr()), C.c_void_p(self.time.data_ptr()), C.c_void_p(self.inst.data_ptr()),
                                         self._stream())
        if rc:
            raise EFError(rc, "ef_op_splat_predict_inst")
        return self.image, self.vertex, self.normal, self.time

    def fill_in(self, depth_mm: torch.Tensor, rgba: torch.Tensor, passthrough=False):
        """FillIn::vertex / normal / image on the last prediction (ElasticFusion.cpp:756-760) -> (image, vertex, normal)"""
        cx, cy, fx, fy = self.cam
        a = (self.height, self.width, C.c_float(cx), C.c_float(cy), C.c_float(fx), C.c_float(fy), int(passthrough))
        rc = self._L.ef_op_fill_vertex(C.c_void_p(self.vertex.data_ptr()), C.c_void_p(depth_mm.data_ptr()), *a,
                                       C.c_void_p(self.filled_vertex.data_ptr()), self._stream())
        # (FillIn::normal keys on the z of the existing NORMAL: Shaders/FillIn.cpp:150-180, fill_normal.frag:44)
        rc = rc or self._L.ef_op_fill_normal(C.c_void_p(self.normal.data_ptr()), C.c_void_p(depth_mm.data_ptr()), *a,
                                             C.c_void_p(self.filled_normal.data_ptr()), self._stream())
        rc = rc or self._L.ef_op_fill_rgb(C.c_void_p(self.image.data_ptr()), C.c_void_p(rgba.data_ptr()), self.height, self.width, int(passthrough),
                                          C.c_void_p(self.filled_image.data_ptr()), self._stream())
        if rc:
            raise EFError(rc, "ef_op_fill_*")
        return self.filled_image, self.filled_vertex, self.filled_normal
