"""time ef_op_depth_bilateral / ef_op_depth_metric with CUDA events (warm), 640x480 and 1280x720"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instancefusion_b200 import binding
from tests.test_depth_prepath import _raw_depth
L = binding.lib()
for w, h in ((640, 480), (1280, 720)):
    d = _raw_depth(w, h)
    src = torch.from_numpy(d.view(np.int16)).cuda()
    dst = torch.zeros_like(src); dstf = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def run(n, metric=False):
        for _ in range(n):
            if metric: L.ef_op_depth_metric(C.c_void_p(src.data_ptr()), C.c_size_t(0), h, w, C.c_float(4.0), C.c_void_p(dstf.data_ptr()), C.c_size_t(0), st)
            else: L.ef_op_depth_bilateral(C.c_void_p(src.data_ptr()), C.c_size_t(0), h, w, C.c_float(4.0), C.c_void_p(dst.data_ptr()), C.c_size_t(0), st)
    for metric in (False, True):
        run(10, metric); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(200, metric); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        valid = int(((d <= 4000) & (d >= 300)).sum())
        if metric: print(f"{w}x{h} depth_metric    {us:7.2f} us  ({w*h*6/us/1e3:.0f} GB/s of 6 B/px)")
        else: print(f"{w}x{h} depth_bilateral {us:7.2f} us  ({valid*169/us/1e3:.1f} G taps/s, {valid} valid px x 169 taps)")
