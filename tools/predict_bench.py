"""Device time of the CUDA model prediction (ef_op_splat_predict) and fill-in: surfel maps made of 1 .. 8 synthetic frames
(one surfel per valid pixel and frame, 640x480), drawn from a pose between them.  Developer tool; run on a GPU box."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instancefusion_b200 import binding, synth

w, h = 640, 480
K = synth.Intrinsics.kinect(w, h)
poses = synth.trajectory(9, seed=2024)
L = binding.lib()
maps = []
for k in range(8):
    f = synth.render(poses[k], K, seed=2024, frame_id=k)
    maps.append(synth.surfels_from_frame(poses[k].numpy(), f["vmap"].numpy(), f["nmap"].numpy(), f["rgba"].numpy(), K, time=k + 1))
for stride, nframes in ((12, 1), (12, 4), (12, 8), (64, 8)):
    s = np.concatenate(maps[:nframes])
    if stride != 12:
        s = np.concatenate([s, np.zeros((len(s), stride - 12), np.float32)], 1)
    d = torch.from_numpy(s).cuda()
    t_inv = np.ascontiguousarray(np.linalg.inv(poses[nframes // 2].numpy().astype(np.float64)).astype(np.float32).reshape(16))
    keys = torch.empty(L.ef_op_splat_scratch_bytes(h, w), dtype=torch.uint8, device="cuda")
    img = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
    v = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    n = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    tm = torch.empty((h, w), dtype=torch.int16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def run():
        rc = L.ef_op_splat_predict(C.c_void_p(d.data_ptr()), C.c_size_t(stride * 4), len(s), t_inv.ctypes.data_as(C.c_void_p), C.c_float(K.cx),
                                   C.c_float(K.cy), C.c_float(K.fx), C.c_float(K.fy), h, w, C.c_float(20.0), C.c_float(9.0), 9, 9, 200,
                                   C.c_void_p(keys.data_ptr()), C.c_void_p(img.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(n.data_ptr()),
                                   C.c_void_p(tm.data_ptr()), st)
        assert rc == 0
    for _ in range(5): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    cover = float((v[..., 2] > 0).float().mean())
    print(f"surfels {len(s):8d} stride {stride*4:3d} B: {ms*1e3:7.1f} us per prediction  ({len(s)*48/ms/1e6:7.1f} GB/s of surfel payload, coverage {cover:.3f})")
