#!/usr/bin/env python
"""bench.py -- joint ICP+RGB dense tracking throughput on B200 (BASELINE.json metric).

A "step" is one tracked frame of the synthetic handheld trajectory (config 2 of BASELINE.json:
640x480, joint ICP+RGB, icpWeight=10, iterations {10,5,4}), i.e. the frameToModel call sequence of
ElasticFusion::processFrame (ElasticFusion.cpp:343-368):
    initICPModel(model vertices, normals, 20 m, pose k-1); initRGBModel(model image);
    initICP(depth k, 20 m); initRGB(rgb k); getIncrementalTransformation(GT pose k-1, ...)
Open-loop protocol: the model maps for frame k are ray-cast at the ground-truth pose k-1.

  value  frames/s with every input already resident in HBM (device pointers into a pre-rendered sequence
         larger than L2), one tracker handle, blocking API -- timed with CUDA events on the handle's stream.
  e2e    frames/s through the host-buffer entry points of the C ABI (ef_init_*_host): every step copies
         its 12.9 MB of inputs from pinned host memory and reads the pose + stats back.  Three full-GPU
         handles take turns so that the copies of two overlap the solve of the third (the copies are the
         bound: 54.6 GB/s measured -> 4 229 frames/s).
  roofline  the persistent tracker kernel (all 19 Gauss-Newton iterations of a frame in one launch):
         algorithmic bytes (ICP 48 B/px/iter + RGB 28 B/px/iter) / its CUDA-event duration
         against the measured HBM copy bandwidth.
  cpu_baseline  the OpenMP C restatement (oracle/, test infrastructure) on a bounded sample of the same
         frames, all host cores.
  --impl reference   the reference's OWN CUDA kernels (oracle/_ref/libef_ref.so built unmodified from
         /root/reference) under the restated host loop, same frames, inputs resident.

Multi-GPU (torchrun): the path does not shard -- every rank tracks its own independent
sequence on its own GPU, no data-path collective; value = total frames / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--frames", type=int, default=301, help="distinct frames rendered into HBM")
    ap.add_argument("--solve", choices=["device", "host"], default="device")
    ap.add_argument("--so3", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0, help="--solve host: replay each iteration's launches from a CUDA graph (EF_OPT_USE_GRAPH)")
    ap.add_argument("--icp-weight", type=float, default=10.0)
    ap.add_argument("--e2e-frames", type=int, default=48, help="distinct frames kept in pinned host memory for e2e")
    ap.add_argument("--cpu-sample", type=int, default=40, help="frames of the CPU baseline sample (0 = skip)")
    ap.add_argument("--ref-kind", choices=["cuda", "port"], default="cuda")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--inflight", type=int, default=4, help="handles (frames in flight) of the pipelined / e2e runs")
    ap.add_argument("--handle-ctas", type=int, default=-1,
                    help="resident pipelined run: CTAs (SMs) each in-flight handle's tracker kernel occupies: -1 = SMs / inflight "
                         "(disjoint SM subsets), 0 = every SM")
    ap.add_argument("--e2e-inflight", type=int, default=3,
                    help="handles of the e2e run; each uses EVERY SM, so their tracker kernels take turns while the other "
                         "handles' host->device copies run (measured best: 3; the copies are the bound)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def u16_np(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


# ------------------------------------------------------------------------------------------------
def render_sequence(args, seed, device):
    """frames rendered on the GPU straight into HBM-resident tensors."""
    from instancefusion_b200 import synth
    K = synth.Intrinsics.kinect(args.width, args.height)
    poses = synth.trajectory(args.frames, seed=seed)
    depth, rgba, vmap, nmap = [], [], [], []
    for k in range(args.frames):
        f = synth.render(poses[k], K, seed=seed, frame_id=k, device=device)
        depth.append(f["depth"])
        rgba.append(f["rgba"])
        vmap.append(f["vmap"])
        nmap.append(f["nmap"])
    torch.cuda.synchronize()
    return K, poses.numpy(), depth, rgba, vmap, nmap


def algorithmic_bytes_per_frame(args):
    """icpStep 48 B/px/iter, RGB pair 28 B/px/iter, so3Step 2 B/px/iter (level 2)."""
    iters = [10, 5, 4]
    n = [(args.width >> i) * (args.height >> i) for i in range(3)]
    weighted = sum(i * p for i, p in zip(iters, n))
    icp = args.icp_weight > 0
    rgb = args.icp_weight < 100
    b = weighted * ((48 if icp else 0) + (28 if rgb else 0))
    return b, weighted


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, device):
    import instancefusion_b200 as ef
    from instancefusion_b200 import rgbd_odometry as RO

    seed = 2024 + rank
    K, poses, depth, rgba, vmap, nmap = render_sequence(args, seed, device)
    F = args.frames
    posef = poses.astype(np.float32)
    mode = RO.EF_SOLVE_DEVICE if args.solve == "device" else RO.EF_SOLVE_HOST
    so3 = bool(args.so3)

    def make():
        tr = ef.RGBDOdometry(args.width, args.height, K.cx, K.cy, K.fx, K.fy, solve_mode=mode)
        if args.graph:
            tr.set_option(RO.EF_OPT_USE_GRAPH, 1)
        if os.environ.get("EF_FRAME_BUILD"):  # A/B switch for experiments: EF_OPT_FRAME_BUILD 0 / 1 / 2
            tr.set_option(RO.EF_OPT_FRAME_BUILD, int(os.environ["EF_FRAME_BUILD"]))
        return tr

    tr = make()
    stream = torch.cuda.ExternalStream(tr.stream)

    def step_resident(i):
        k = 1 + (i % (F - 1))
        # = initICPModel, initRGBModel, initICP, initRGB, getIncrementalTransformation (ef_track_frame_to_model)
        return tr.trackFrameToModel(vmap[k - 1], nmap[k - 1], rgba[k - 1], depth[k], rgba[k], 20.0, posef[k - 1], False, args.icp_weight,
                                    True, False, so3), k

    if so3:
        tr.initFirstRGB(rgba[0])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: resident inputs ----
    for i in range(args.warmup):
        step_resident(i)
    errs = []
    barrier()
    clocks = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    clocks.start()
    launches0 = tr.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    results = []
    for i in range(args.steps):
        results.append(step_resident(args.warmup + i))  # the pose is on the host when the call returns
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    errs = [float(np.linalg.norm(t - poses[k][:3, 3])) for (t, R), k in results]  # scored against the ground truth outside the timed region
    ev_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = tr.launch_count - launches0
    ms_total = max(ev_ms, 1e-9)

    # ---- roofline of the dominant kernel: second pass with CUDA events around each solve ----
    tr.set_option(RO.EF_OPT_PROFILE, 1)
    tr.profile()
    for i in range(args.steps):
        step_resident(args.warmup + i)
    solve_ms, calls = tr.profile()
    tr.set_option(RO.EF_OPT_PROFILE, 0)

    # ---- the same frames through the reference's own five calls (initICPModel, initRGBModel, initICP, initRGB,
    #      getIncrementalTransformation: ElasticFusion.cpp:343-368) instead of the single-call entry; single rank only ----
    five_call = None
    if world == 1:
        def step_five(i):
            k = 1 + (i % (F - 1))
            tr.initICPModel(vmap[k - 1], nmap[k - 1], 20.0, posef[k - 1])
            tr.initRGBModel(rgba[k - 1])
            tr.initICP(depth[k], 20.0)
            tr.initRGB(rgba[k])
            return tr.getIncrementalTransformation(posef[k - 1][:3, 3], posef[k - 1][:3, :3], False, args.icp_weight, True, False, so3)
        n5 = max(1, min(args.steps, 200))
        for i in range(min(args.warmup, 10)):
            step_five(i)
        torch.cuda.synchronize()
        l0 = tr.launch_count
        t5 = time.perf_counter()
        for i in range(n5):
            step_five(args.warmup + i)
        torch.cuda.synchronize()
        five_call = {"value": n5 / (time.perf_counter() - t5), "unit": "frames/s", "launches_per_frame": (tr.launch_count - l0) / n5,
                     "note": "the five calls of the reference's frameToModel sequence (RGBDOdometry API) instead of ef_track_frame_to_model"}
        # ... and with EF_OPT_DEFER_BUILD: the init* calls record their arguments, one builder launch at getIncrementalTransformation
        tr.set_option(RO.EF_OPT_DEFER_BUILD, 1)
        for i in range(min(args.warmup, 10)):
            step_five(i)
        torch.cuda.synchronize()
        l0 = tr.launch_count
        t5 = time.perf_counter()
        for i in range(n5):
            step_five(args.warmup + i)
        torch.cuda.synchronize()
        five_call["deferred_build"] = {"value": n5 / (time.perf_counter() - t5), "launches_per_frame": (tr.launch_count - l0) / n5}
        tr.set_option(RO.EF_OPT_DEFER_BUILD, 0)

    # ---- pipelined runs: `inflight` handles, each on its own share of the SMs (EF_OPT_GRID_CTAS), track
    #      consecutive frames concurrently (the open-loop protocol makes frames independent).  The solve of one
    #      frame is a chain of L2 round trips, so frames overlap almost perfectly; with host buffers the H2D copies
    #      of one handle overlap the solves of the others. ----
    e2e = None
    concurrent = None
    if not args.no_e2e:
        NH = max(1, args.inflight)
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        share = 0
        if mode == RO.EF_SOLVE_DEVICE and args.handle_ctas != 0:
            # a handle's share of the SMs must be able to hold its photometric candidates in shared memory: halve the
            # number of handles until the tracker accepts the split (1280x720 takes 2 handles, 640x480 takes 4)
            while NH > 1:
                share = args.handle_ctas if args.handle_ctas > 0 else sms // NH
                try:
                    tr.set_option(RO.EF_OPT_GRID_CTAS, share)
                    break
                except Exception:
                    NH //= 2
            if NH == 1:
                share = 0
        if mode == RO.EF_SOLVE_DEVICE:
            tr.set_option(RO.EF_OPT_GRID_CTAS, share)
        trs = [tr] + [make() for _ in range(NH - 1)]
        if mode == RO.EF_SOLVE_DEVICE:
            for t_ in trs[1:]:
                t_.set_option(RO.EF_OPT_GRID_CTAS, share)
        ctas_per_handle = share if share > 0 else sms
        if so3:
            for t_ in trs[1:]:
                t_.initFirstRGB(rgba[0])

        def pipelined(submit, n):
            for i in range(min(args.warmup, 2 * NH)):
                submit(trs[i % NH], i)
                trs[i % NH].finish()
            barrier()
            t0 = time.perf_counter()
            inflight = []
            for i in range(n):
                trk = trs[i % NH]
                if len(inflight) == NH:
                    inflight.pop(0).finish()
                submit(trk, i)
                inflight.append(trk)
            for trk in inflight:
                trk.finish()
            barrier()
            return time.perf_counter() - t0

        def submit_resident(trk, i):
            k = 1 + (i % (F - 1))
            trk.trackFrameToModelLaunch(vmap[k - 1], nmap[k - 1], rgba[k - 1], depth[k], rgba[k], 20.0, posef[k - 1], False, args.icp_weight,
                                        True, False, so3)

        concurrent = {"seconds": pipelined(submit_resident, args.steps), "frames": args.steps, "handles": NH, "ctas_per_handle": ctas_per_handle}

        # e2e: host buffers.  The 12.9 MB of a frame take 0.24 ms over PCIe (54.6 GB/s measured), about as long as the
        # whole solve, so the best schedule is full-GPU handles taking turns: one computes while the others copy.
        NE = max(1, min(args.e2e_inflight, 8))
        while len(trs) < NE:
            trs.append(make())
            if so3:
                trs[-1].initFirstRGB(rgba[0])
        if mode == RO.EF_SOLVE_DEVICE:
            for t_ in trs:
                t_.set_option(RO.EF_OPT_GRID_CTAS, 0)
        NH_resident = NH
        trs_all, trs, NH = trs, trs[:NE], NE
        FE = min(args.e2e_frames, F)
        pin = lambda t: t.cpu().pin_memory()
        h_depth = [pin(depth[k].view(torch.int16)) for k in range(FE)]
        h_rgba = [pin(rgba[k]) for k in range(FE)]
        h_vmap = [pin(vmap[k]) for k in range(FE)]
        h_nmap = [pin(nmap[k]) for k in range(FE)]

        def submit_host(trk, i):
            k = 1 + (i % (FE - 1))
            trk.trackFrameToModelLaunch(h_vmap[k - 1], h_nmap[k - 1], h_rgba[k - 1], h_depth[k], h_rgba[k], 20.0, posef[k - 1], False,
                                        args.icp_weight, True, False, so3)

        e2e_s = pipelined(submit_host, args.steps)
        h2d = args.width * args.height * (2 + 4 + 16 + 16 + 4)  # depth + rgb + vmap + nmap + model rgb
        e2e = {"seconds": e2e_s, "frames": args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 48 + 384, "handles": NE,
               "ctas_per_handle": sms}
        trs = trs_all
        for t_ in trs[1:]:
            t_.close()

    # ---- cpu baseline (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        from oracle import oracle as O
        n = min(args.cpu_sample, F - 1)
        host = [(u16_np(depth[k]), rgba[k].cpu().numpy(), vmap[k].cpu().numpy(), nmap[k].cpu().numpy()) for k in range(n + 1)]
        cpu = O.OracleTracker(args.width, args.height, K.cx, K.cy, K.fx, K.fy, impl="cpu")
        t0 = time.perf_counter()
        for k in range(1, n + 1):
            p = posef[k - 1]
            cpu.init_icp_model(host[k - 1][2], host[k - 1][3], 20.0, p)
            cpu.init_rgb_model(host[k - 1][1])
            cpu.init_icp_depth(host[k][0], 20.0)
            cpu.init_rgb(host[k][1])
            cpu.get_incremental_transformation(p[:3, 3], p[:3, :3], False, args.icp_weight, True, False, so3)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"first {n} frames of the same trajectory, OpenMP C restatement (oracle/ef_oracle.c), all host cores"}
        cpu.close()

    tr.close()
    return {"ms_total": ms_total, "wall_ms": wall_ms, "launches": launches, "clocks": clk, "solve_ms": solve_ms, "solve_calls": calls,
            "e2e": e2e, "concurrent": concurrent, "five_call": five_call, "cpu_baseline": cpu_baseline, "median_err_m": float(np.median(errs)), "max_err_m": float(np.max(errs))}


def run_reference(args, device):
    """reference CUDA operators (oracle/_ref) under the restated RGBDOdometry host loop; inputs resident."""
    from oracle import oracle as O
    so3 = bool(args.so3)
    K, poses, depth, rgba, vmap, nmap = render_sequence(args, 2024, device)
    F = args.frames
    posef = poses.astype(np.float32)
    if args.ref_kind == "port" or not O.ref_available():
        kind = "port"
        tr = O.OracleTracker(args.width, args.height, K.cx, K.cy, K.fx, K.fy, impl="cpu")
        host = {}

        def step(i):
            k = 1 + (i % (F - 1))
            for j in (k - 1, k):
                if j not in host:
                    host[j] = (u16_np(depth[j]), rgba[j].cpu().numpy(), vmap[j].cpu().numpy(), nmap[j].cpu().numpy())
            p = posef[k - 1]
            tr.init_icp_model(host[k - 1][2], host[k - 1][3], 20.0, p)
            tr.init_rgb_model(host[k - 1][1])
            tr.init_icp_depth(host[k][0], 20.0)
            tr.init_rgb(host[k][1])
            tr.get_incremental_transformation(p[:3, 3], p[:3, :3], False, args.icp_weight, True, False, so3)
    else:
        kind = "reference"
        tr = O.OracleTracker(args.width, args.height, K.cx, K.cy, K.fx, K.fy, impl="ref")
        vp = lambda t: C.c_void_p(t.data_ptr())

        def step(i):
            k = 1 + (i % (F - 1))
            p = posef[k - 1]
            pp = np.ascontiguousarray(p.reshape(16))
            tr.call_dev("init_icp_model", vp(vmap[k - 1]), vp(nmap[k - 1]), C.c_float(20.0), pp.ctypes.data_as(C.c_void_p))
            tr.call_dev("init_rgb_model", vp(rgba[k - 1]))
            tr.call_dev("init_icp_depth", vp(depth[k]), C.c_float(20.0))
            tr.call_dev("init_rgb", vp(rgba[k]))
            tr.get_incremental_transformation(p[:3, 3], p[:3, :3], False, args.icp_weight, True, False, so3)
        if so3:
            tr.call_dev("init_first_rgb", vp(rgba[0]))

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    tr.close()
    return kind, dt, clk


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tracker has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    workload = (f"synthetic {args.width}x{args.height} {args.frames}-frame handheld trajectory, joint ICP+RGB "
                f"(icpWeight={args.icp_weight:g}, iters {{10,5,4}}, so3={args.so3}), open-loop frame-to-model")
    config = {"workload": workload, "solve": args.solve + ("+graph" if args.graph else ""), "frames_resident": args.frames,
              "l2": "each step reads a different frame of a >3 GB resident sequence (inputs larger than the 126 MB L2)",
              "parallelism": f"replicas x{world} (independent sequences, no collective)"}

    if args.impl == "reference":
        if rank != 0:
            return
        kind, dt, clk = run_reference(args, device)
        fps = args.steps / dt
        line = {"impl": "reference", "metric": "joint ICP+RGB tracking frames/s", "value": fps, "unit": "frames/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clk,
                "reference_kind": ("reference CUDA kernels (oracle/_ref/libef_ref.so, built unmodified from the reference sources, "
                                   "GPUConfig default launch shapes) on this GPU" if kind == "reference"
                                   else "OpenMP C restatement of the reference kernels on the host cores"),
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count() if kind == "port" else 1, "kind": kind,
                                 "sample": f"{args.steps} frames of the same trajectory"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)

    r = run_ours(args, rank, world, device)

    ms = torch.tensor([r["ms_total"], r["e2e"]["seconds"] * 1e3 if r["e2e"] else 0.0, float(r["solve_ms"]),
                       r["concurrent"]["seconds"] * 1e3 if r["concurrent"] else 0.0], device=device, dtype=torch.float64)
    launches = torch.tensor([float(r["launches"])], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(launches, op=torch.distributed.ReduceOp.SUM)
    ms_total, e2e_ms, solve_ms, conc_ms = [float(x) for x in ms.tolist()]

    if rank == 0:
        total_frames = args.steps * world
        value = total_frames / (ms_total * 1e-3)
        peak, peak_src = measured_peaks()
        bytes_frame, weighted_px = algorithmic_bytes_per_frame(args)
        avg_solve_ms = solve_ms / max(r["solve_calls"], 1)
        achieved = bytes_frame / (avg_solve_ms * 1e-3) / 1e9 if avg_solve_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"{args.width}x{args.height}")
            except Exception:
                traffic = None
        line = {
            "metric": "joint ICP+RGB tracking frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": r["clocks"],
            "gpu_launches": int(launches.item()),
            "us_per_gn_iteration": avg_solve_ms * 1e3 / 19.0,
            "tracking_error_m": {"median": r["median_err_m"], "max": r["max_err_m"]},
            "roofline": {"bound": "hbm", "kernel": "k_track (persistent tracker kernel: 19 Gauss-Newton iterations per launch)"
                         if args.solve == "device" else "step loop (icp/rgbres/rgb kernels, host solve)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_frame,
                         "avg_launch_ms": avg_solve_ms,
                         "note": "48 B/px/iter ICP + 28 B/px/iter RGB over 10*N0+5*N1+4*N2 pixels; the working set is L2-resident "
                                 "after the first iteration, so the bound that binds is the per-iteration grid barrier + solve latency"},
        }
        if r["e2e"]:
            line["e2e"] = {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s",
                           "h2d_bytes_per_step": r["e2e"]["h2d_bytes_per_step"], "d2h_bytes_per_step": r["e2e"]["d2h_bytes_per_step"],
                           "inflight_frames": r["e2e"]["handles"], "ctas_per_handle": r["e2e"]["ctas_per_handle"],
                           "note": "C ABI ef_track_frame_to_model with pinned host buffers; full-GPU handles take turns: one solves while "
                                   "the others copy (PCIe bound: 12.9 MB per frame at 640x480)"}
            line["value_pipelined"] = {"value": total_frames / (conc_ms * 1e-3), "unit": "frames/s", "handles": r["concurrent"]["handles"],
                                       "ctas_per_handle": r["concurrent"]["ctas_per_handle"],
                                       "note": "inputs resident, handles on disjoint SM subsets: throughput when consecutive frames may "
                                               "overlap (`value` is the single-handle, one-frame-at-a-time rate)"}
        if r.get("five_call"):
            line["value_reference_api"] = r["five_call"]
        if r["cpu_baseline"]:
            line["cpu_baseline"] = r["cpu_baseline"]
        print(json.dumps(line))

    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
