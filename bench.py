#!/usr/bin/env python
"""bench.py -- joint ICP+RGB dense tracking throughput on B200 (BASELINE.json metric).

A "step" is one tracked frame of the synthetic handheld trajectory (config 2 of BASELINE.json:
640x480, joint ICP+RGB, icpWeight=10, iterations {10,5,4}), i.e. the frameToModel call sequence of
ElasticFusion::processFrame (ElasticFusion.cpp:343-368):
    initICPModel(model vertices, normals, 20 m, pose k-1); initRGBModel(model image);
    initICP(depth k, 20 m); initRGB(rgb k); getIncrementalTransformation(GT pose k-1, ...)
Open-loop protocol: the model maps for frame k are ray-cast at the ground-truth pose k-1.

  value    frames/s with every input already resident in HBM (device pointers into a pre-rendered sequence larger than
           L2), one tracker handle, one frame at a time, blocking single-call entry ef_track_frame_to_model -- CUDA events
           on the handle's stream.
  value_reference_api   the same frames through the reference's own FIVE calls (class RGBDOdometry's API, what
           include/compat/RGBDOdometry.h forwards): THE DROP-IN NUMBER.  `deferred_build` = with EF_OPT_DEFER_BUILD (the
           shim's setting: 2 launches per frame).
  value_1280x720  BASELINE configs[2]: the same trajectory at 1280x720 with SO(3) pre-alignment, with its own roofline.
  e2e      frames/s through the host-buffer entry points of the C ABI: every step copies its 12.9 MB of inputs from
           pinned host memory and reads the pose + stats back.  `value` = three full-GPU handles taking turns (open-loop
           frames are independent: the copies of two overlap the solve of the third); `single` = ONE frame in flight, the
           rate a closed-loop caller sees; `sensor_only` = the production data flow: the model maps never cross PCIe (GL
           textures in the reference), they are predicted on the device from a resident surfel map (ef_op_splat_predict,
           inside the timed region) and only the 1.8 MB sensor frame is copied.
  roofline the persistent tracker kernel (all 19 Gauss-Newton iterations of a frame in one launch): algorithmic bytes
           (ICP 48 B/px/iter + RGB 28 B/px/iter) / its CUDA-event duration against the measured HBM copy
           bandwidth; `levels` = the same per pyramid level (share of the launch from the kernel's clock64 trace);
           `traffic` / `l2_bytes` from the committed ncu capture (profiles/roofline_traffic.json).
  cpu_baseline  the OpenMP C restatement (oracle/, test infrastructure) on a bounded sample of the same frames.
  --impl reference   the reference's OWN CUDA kernels (oracle/_ref/libef_ref.so built unmodified from /root/reference)
           under the restated host loop, same frames, inputs resident, one replica per rank: `value` at GPUConfig's default
           launch shapes (the stock path), `value_swept` at the best shapes of a GPUTest-style sweep (GPUTest.cpp:247-324).

Multi-GPU (torchrun): the path does not shard -- every rank tracks its own independent sequence on its
own GPU, no data-path collective; value = total frames / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
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
    ap.add_argument("--ref-sweep", type=int, default=1, help="--impl reference: also time the best launch shapes of a GPUTest-style sweep")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-720p", action="store_true", help="skip the 1280x720 + SO(3) workload (configs[2]) of the N=1 line")
    ap.add_argument("--no-levels", action="store_true", help="skip the per-level roofline pass (clock64 trace)")
    ap.add_argument("--inflight", type=int, default=4, help="handles (frames in flight) of the pipelined run")
    ap.add_argument("--handle-ctas", type=int, default=-1,
                    help="resident pipelined run: CTAs (SMs) each in-flight handle's tracker kernel occupies: -1 = SMs / inflight "
                         "(disjoint SM subsets), 0 = every SM")
    ap.add_argument("--e2e-inflight", type=int, default=3,
                    help="handles of the e2e run; each uses EVERY SM, so their tracker kernels take turns while the other "
                         "handles' host->device copies run (measured best: 3; the copies are the bound)")
    ap.add_argument("--keyframe", type=int, default=30, help="sensor-only e2e: the resident surfel map is re-seeded every this many frames (outside the timed region)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons of THIS rank's GPU during the timed region (B200_PROFILING.md), sampled through NVML in a
    thread (one nvidia-smi process per rank produced no samples when 4-8 ranks started it at once)."""

    def __init__(self, device):
        self.samples, self.reasons, self.mx = [], set(), None
        self.stop_flag = False
        self.thread = None
        self.err = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                idx = device.index if device.index is not None else 0
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                if vis:
                    idx = int(vis.split(",")[idx])
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        time.sleep(0.03)  # at least a couple of samples even for a very short region
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["no samples: " + str(self.err)]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.samples)}


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
def render_sequence(width, height, frames, seed, device):
    """frames rendered on the GPU straight into HBM-resident tensors."""
    from instancefusion_b200 import synth
    K = synth.Intrinsics.kinect(width, height)
    poses = synth.trajectory(frames, seed=seed)
    depth, rgba, vmap, nmap = [], [], [], []
    for k in range(frames):
        f = synth.render(poses[k], K, seed=seed, frame_id=k, device=device)
        depth.append(f["depth"])
        rgba.append(f["rgba"])
        vmap.append(f["vmap"])
        nmap.append(f["nmap"])
    torch.cuda.synchronize()
    return K, poses.numpy(), depth, rgba, vmap, nmap


ITERS = [10, 5, 4]


def algorithmic_bytes(width, height, icp_weight):
    """icpStep 48 B/px/iter, RGB pair 28 B/px/iter -> (bytes per frame, bytes per level)."""
    n = [(width >> i) * (height >> i) for i in range(3)]
    per_px = (48 if icp_weight > 0 else 0) + (28 if icp_weight < 100 else 0)
    lv = [ITERS[i] * n[i] * per_px for i in range(3)]
    return sum(lv), lv


def roofline_block(width, height, icp_weight, avg_launch_ms, solve, level_share=None):
    peak, peak_src = measured_peaks()
    total, per_level = algorithmic_bytes(width, height, icp_weight)
    achieved = total / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    traffic, l2 = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            rec = json.load(open(tp)).get(f"{width}x{height}")
            if isinstance(rec, dict):
                traffic, l2 = rec.get("dram_bytes"), rec.get("lts_bytes")
            else:
                traffic = rec
        except Exception:
            pass
    out = {"bound": "hbm",
           "kernel": "k_track (persistent tracker kernel: 19 Gauss-Newton iterations per launch)" if solve == "device"
           else "step loop (icp/rgbres/rgb kernels, host solve)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "l2_bytes": l2,
           "peak_source": peak_src, "algorithmic_bytes_per_launch": total, "avg_launch_ms": avg_launch_ms,
           "note": "48 B/px/iter ICP + 28 B/px/iter RGB over 10*N0+5*N1+4*N2 pixels; the working set is L2-resident after the "
                   "first iteration (traffic = DRAM bytes, l2_bytes = lts__t_bytes of one launch, ncu), so what binds is the "
                   "per-iteration solve chain + instruction issue"}
    if level_share:
        out["levels"] = []
        for lv in (2, 1, 0):
            ms = avg_launch_ms * level_share[lv]
            gbs = per_level[lv] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            out["levels"].append({"level": lv, "iterations": ITERS[lv], "ms": ms, "us_per_iteration": ms * 1e3 / ITERS[lv],
                                  "algorithmic_bytes": per_level[lv], "achieved": gbs, "frac": gbs / peak})
    return out


# ------------------------------------------------------------------------------------------------
class Workload:
    """one resolution of the trajectory resident in HBM + the measurement passes over it"""

    def __init__(self, args, width, height, so3, frames, seed, device, world):
        import instancefusion_b200 as ef
        from instancefusion_b200 import rgbd_odometry as RO
        self.ef, self.RO = ef, RO
        self.args, self.w, self.h, self.so3, self.F, self.device, self.world = args, width, height, bool(so3), frames, device, world
        self.K, self.poses, self.depth, self.rgba, self.vmap, self.nmap = render_sequence(width, height, frames, seed, device)
        self.posef = self.poses.astype(np.float32)
        self.mode = RO.EF_SOLVE_DEVICE if args.solve == "device" else RO.EF_SOLVE_HOST
        self.icpw = args.icp_weight

    def make(self):
        K = self.K
        tr = self.ef.RGBDOdometry(self.w, self.h, K.cx, K.cy, K.fx, K.fy, solve_mode=self.mode)
        if self.args.graph:
            tr.set_option(self.RO.EF_OPT_USE_GRAPH, 1)
        if os.environ.get("EF_FRAME_BUILD"):  # A/B switch for experiments: EF_OPT_FRAME_BUILD 0 / 1 / 2
            tr.set_option(self.RO.EF_OPT_FRAME_BUILD, int(os.environ["EF_FRAME_BUILD"]))
        if self.so3:
            tr.initFirstRGB(self.rgba[0])
        return tr

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident(self, tr, i):
        k = 1 + (i % (self.F - 1))
        # = initICPModel, initRGBModel, initICP, initRGB, getIncrementalTransformation (ef_track_frame_to_model)
        return tr.trackFrameToModel(self.vmap[k - 1], self.nmap[k - 1], self.rgba[k - 1], self.depth[k], self.rgba[k], 20.0, self.posef[k - 1],
                                    False, self.icpw, True, False, self.so3), k

    def resident(self, tr, steps, warmup, sample_clocks=True):
        """-> dict(ms_total (CUDA events on the handle's stream), launches, clocks, errs)"""
        stream = torch.cuda.ExternalStream(tr.stream)
        for i in range(warmup):
            self.step_resident(tr, i)
        self.barrier()
        clocks = ClockSampler(self.device) if sample_clocks else None
        if clocks:
            clocks.start()
        l0 = tr.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        results = [self.step_resident(tr, warmup + i) for i in range(steps)]  # the pose is on the host when the call returns
        e1.record(stream)
        self.barrier()
        ms = max(e0.elapsed_time(e1), 1e-9)
        clk = clocks.stop() if clocks else None
        errs = [float(np.linalg.norm(t - self.poses[k][:3, 3])) for (t, R), k in results]  # scored outside the timed region
        return {"ms_total": ms, "launches": tr.launch_count - l0, "clocks": clk, "median_err_m": float(np.median(errs)), "max_err_m": float(np.max(errs))}

    def solve_time(self, tr, steps, warmup):
        """second pass with CUDA events around each solve: average duration of the dominant kernel"""
        tr.set_option(self.RO.EF_OPT_PROFILE, 1)
        tr.profile()
        for i in range(steps):
            self.step_resident(tr, warmup + i)
        ms, calls = tr.profile()
        tr.set_option(self.RO.EF_OPT_PROFILE, 0)
        return ms, calls

    def level_share(self, steps=60):
        """share of a tracker-kernel launch spent at each pyramid level, from the kernel's own clock64 trace (a handle
        created with EF_TRACK_TIMING=1 runs the stamped build of the kernel; only the SHARES are used)"""
        if self.mode != self.RO.EF_SOLVE_DEVICE:
            return None
        os.environ["EF_TRACK_TIMING"] = "1"
        try:
            tr = self.make()
        finally:
            del os.environ["EF_TRACK_TIMING"]
        try:
            for i in range(steps):
                self.step_resident(tr, i)
            cyc = (C.c_double * 32)()
            n = C.c_longlong(0)
            if tr._L.ef_tracker_trace(tr._h, cyc, C.byref(n)) != 0 or n.value == 0:
                return None
            c = [cyc[i] / n.value for i in range(32)]
            end = c[31]
            first = {2: 0, 1: ITERS[2], 0: ITERS[2] + ITERS[1]}
            if end <= 0 or c[first[1]] <= 0 or c[first[0]] <= 0:
                return None
            # level 2 = kernel start (incl. the SO(3) pre-alignment and every level's preparation that is not hidden) .. first
            # iteration of level 1, and so on
            return {2: c[first[1]] / end, 1: (c[first[0]] - c[first[1]]) / end, 0: (end - c[first[0]]) / end}
        finally:
            tr.close()

    def five_call(self, tr, steps, warmup):
        def step(i):
            k = 1 + (i % (self.F - 1))
            tr.initICPModel(self.vmap[k - 1], self.nmap[k - 1], 20.0, self.posef[k - 1])
            tr.initRGBModel(self.rgba[k - 1])
            tr.initICP(self.depth[k], 20.0)
            tr.initRGB(self.rgba[k])
            return tr.getIncrementalTransformation(self.posef[k - 1][:3, 3], self.posef[k - 1][:3, :3], False, self.icpw, True, False, self.so3)

        def timed(n):
            for i in range(min(warmup, 10)):
                step(i)
            torch.cuda.synchronize()
            l0 = tr.launch_count
            t0 = time.perf_counter()
            for i in range(n):
                step(warmup + i)
            torch.cuda.synchronize()
            return n / (time.perf_counter() - t0), (tr.launch_count - l0) / n

        n5 = max(1, min(steps, 200))
        v, l = timed(n5)
        out = {"value": v, "unit": "frames/s", "launches_per_frame": l,
               "note": "the five calls of the reference's frameToModel sequence (class RGBDOdometry's API) instead of ef_track_frame_to_model: "
                       "the drop-in number"}
        tr.set_option(self.RO.EF_OPT_DEFER_BUILD, 1)
        v, l = timed(n5)
        out["deferred_build"] = {"value": v, "launches_per_frame": l,
                                 "note": "EF_OPT_DEFER_BUILD, the setting of include/compat/RGBDOdometry.h: init* calls record, one builder launch at the solve"}
        tr.set_option(self.RO.EF_OPT_DEFER_BUILD, 0)
        return out

    # ---- pipelined runs: several handles, frames in flight ----
    def pipelined(self, trs, submit, n, warmup):
        NH = len(trs)
        for i in range(min(warmup, 2 * NH)):
            submit(trs[i % NH], i)
            trs[i % NH].finish()
        self.barrier()
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
        self.barrier()
        return time.perf_counter() - t0


def run_ours(args, rank, world, device):
    from instancefusion_b200 import rgbd_odometry as RO
    seed = 2024 + rank
    W = Workload(args, args.width, args.height, args.so3, args.frames, seed, device, world)
    tr = W.make()
    sms = torch.cuda.get_device_properties(device).multi_processor_count

    # ---- value: resident inputs, one handle, one frame at a time ----
    r = W.resident(tr, args.steps, args.warmup)
    solve_ms, calls = W.solve_time(tr, args.steps, args.warmup)
    out = {"ms_total": r["ms_total"], "launches": r["launches"], "clocks": r["clocks"], "solve_ms": solve_ms, "solve_calls": calls,
           "median_err_m": r["median_err_m"], "max_err_m": r["max_err_m"], "e2e": None, "concurrent": None, "five_call": None,
           "cpu_baseline": None, "level_share": None, "w720": None, "e2e_single_s": 0.0, "sensor": None, "batched": None}
    if rank == 0 and not args.no_levels:
        out["level_share"] = W.level_share()

    # ---- the same frames through the reference's own five calls; single rank only ----
    if world == 1:
        out["five_call"] = W.five_call(tr, args.steps, args.warmup)

    if not args.no_e2e:
        F, so3, mode = W.F, W.so3, W.mode
        # resident, `inflight` handles on disjoint SM subsets (EF_OPT_GRID_CTAS): consecutive frames overlap
        NH = max(1, args.inflight)
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
        trs = [tr] + [W.make() for _ in range(NH - 1)]
        if mode == RO.EF_SOLVE_DEVICE:
            for t_ in trs[1:]:
                t_.set_option(RO.EF_OPT_GRID_CTAS, share)

        def submit_resident(trk, i):
            k = 1 + (i % (F - 1))
            trk.trackFrameToModelLaunch(W.vmap[k - 1], W.nmap[k - 1], W.rgba[k - 1], W.depth[k], W.rgba[k], 20.0, W.posef[k - 1], False, W.icpw,
                                        True, False, so3)

        out["concurrent"] = {"seconds": W.pipelined(trs, submit_resident, args.steps, args.warmup), "handles": NH,
                             "ctas_per_handle": share if share > 0 else sms}

        # k sequences per launch (ef_track_frames_to_model_batch): two handles track two sequences -- the same trajectory half
        # a lap apart -- with ONE persistent kernel per frame pair on every SM; blocking call, one pair at a time
        if mode == RO.EF_SOLVE_DEVICE:
            try:
                for t_ in trs[:2]:
                    t_.set_option(RO.EF_OPT_GRID_CTAS, 0)
                while len(trs) < 2:
                    trs.append(W.make())
                bt = RO.BatchTracker(trs[:2])
                off = (F - 1) // 2

                def pair(i):
                    ks = [1 + (i % (F - 1)), 1 + ((i + off) % (F - 1))]
                    fr = [(W.vmap[k - 1], W.nmap[k - 1], W.rgba[k - 1], W.depth[k], W.rgba[k]) for k in ks]
                    return bt.track(fr, [W.posef[k - 1] for k in ks], 20.0, False, W.icpw, True, False, so3), ks

                npairs = max(1, args.steps // 2)
                for i in range(min(args.warmup, 10)):
                    pair(i)
                W.barrier()
                st = torch.cuda.ExternalStream(trs[0].stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                res = [pair(args.warmup + i) for i in range(npairs)]
                e1.record(st)
                W.barrier()
                errs = [float(np.linalg.norm(t - W.poses[k][:3, 3])) for r_, ks in res for (t, R), k in zip(r_, ks)]
                out["batched"] = {"ms_total": max(e0.elapsed_time(e1), 1e-9), "frames": 2 * npairs, "sequences": 2, "median_err_m": float(np.median(errs))}
                # the same with two batches in flight: a second pair of handles is launched before the first is finished
                while len(trs) < 4:
                    trs.append(W.make())
                for t_ in trs[2:4]:
                    t_.set_option(RO.EF_OPT_GRID_CTAS, 0)
                bts = [bt, RO.BatchTracker(trs[2:4])]

                def launch_pair(b, i):
                    ks = [1 + (i % (F - 1)), 1 + ((i + off) % (F - 1))]
                    fr = [(W.vmap[k - 1], W.nmap[k - 1], W.rgba[k - 1], W.depth[k], W.rgba[k]) for k in ks]
                    b.launch(fr, [W.posef[k - 1] for k in ks], 20.0, False, W.icpw, True, False, so3)

                for i in range(4):
                    launch_pair(bts[i % 2], i)
                    bts[i % 2].finish()
                W.barrier()
                t0 = time.perf_counter()
                launch_pair(bts[0], 0)
                for i in range(1, npairs):
                    launch_pair(bts[i % 2], i)
                    bts[(i - 1) % 2].finish()
                bts[(npairs - 1) % 2].finish()
                W.barrier()
                out["batched"]["inflight2_ms_total"] = (time.perf_counter() - t0) * 1e3
                if mode == RO.EF_SOLVE_DEVICE:
                    for t_ in trs[:2]:
                        t_.set_option(RO.EF_OPT_GRID_CTAS, share)
            except Exception as e:  # noqa: BLE001
                out["batched"] = {"error": repr(e)}

        # e2e: host buffers.  The 12.9 MB of a frame take 0.24 ms over PCIe, about as long as the whole solve, so the best
        # schedule is full-GPU handles taking turns: one computes while the others copy.
        NE = max(1, min(args.e2e_inflight, 8))
        while len(trs) < NE:
            trs.append(W.make())
        if mode == RO.EF_SOLVE_DEVICE:
            for t_ in trs:
                t_.set_option(RO.EF_OPT_GRID_CTAS, 0)
        FE = min(args.e2e_frames, F)
        pin = lambda t: t.cpu().pin_memory()  # noqa: E731
        h_depth = [pin(W.depth[k].view(torch.int16)) for k in range(FE)]
        h_rgba = [pin(W.rgba[k]) for k in range(FE)]
        h_vmap = [pin(W.vmap[k]) for k in range(FE)]
        h_nmap = [pin(W.nmap[k]) for k in range(FE)]

        def submit_host(trk, i):
            k = 1 + (i % (FE - 1))
            trk.trackFrameToModelLaunch(h_vmap[k - 1], h_nmap[k - 1], h_rgba[k - 1], h_depth[k], h_rgba[k], 20.0, W.posef[k - 1], False,
                                        W.icpw, True, False, so3)

        e2e_s = W.pipelined(trs[:NE], submit_host, args.steps, args.warmup)
        single_s = W.pipelined(trs[:1], submit_host, args.steps, args.warmup)  # ONE frame in flight: submit, wait, submit, ...
        npx = W.w * W.h
        out["e2e"] = {"seconds": e2e_s, "single_seconds": single_s, "h2d_bytes_per_step": npx * (2 + 4 + 16 + 16 + 4), "d2h_bytes_per_step": 48 + 384,
                      "handles": NE, "ctas_per_handle": sms}

        # e2e, production data flow: surfel map resident on the device, model maps predicted there (ef_op_splat_predict)
        # inside the timed region, only depth + colour (6 B/px) cross PCIe
        try:
            from instancefusion_b200 import synth
            from instancefusion_b200.predict import ModelPredictor
            KF = max(1, args.keyframe)
            keyframes = {}
            for k0 in range(0, FE, KF):  # outside the timed region: one surfel map per key frame, seeded like GlobalModel::initialise
                s = synth.surfels_from_frame(W.poses[k0], h_vmap[k0].numpy(), h_nmap[k0].numpy(), h_rgba[k0].numpy(), W.K, time=1)
                keyframes[k0] = torch.from_numpy(s).to(device)
            preds = []
            for t_ in trs[:NE]:
                p = ModelPredictor(W.w, W.h, W.K.cx, W.K.cy, W.K.fx, W.K.fy)
                p.stream = t_.stream  # prediction and tracking of a handle are ordered on the handle's stream
                preds.append(p)
            pred_of = {id(t_): p for t_, p in zip(trs[:NE], preds)}
            sensor_errs = []

            def submit_sensor(trk, i):
                k = 1 + (i % (FE - 1))
                p = pred_of[id(trk)]
                img, v, nrm, _ = p.predict(keyframes[((k - 1) // KF) * KF], W.poses[k - 1], time=2, maxTime=2, timeDelta=10 ** 6, maxDepth=20.0,
                                           confThreshold=9.0)
                trk.trackFrameToModelLaunch(v, nrm, img, h_depth[k], h_rgba[k], 20.0, W.posef[k - 1], False, W.icpw, True, False, so3)

            s_multi = W.pipelined(trs[:NE], submit_sensor, args.steps, args.warmup)
            s_single = W.pipelined(trs[:1], submit_sensor, args.steps, args.warmup)
            # the same flow with the handles on disjoint SM subsets (EF_OPT_GRID_CTAS, like value_pipelined): PCIe carries only
            # 1.8 MB per frame here, so the GPU is the bound and partitions fill it better than full-GPU handles taking turns
            s_part, n_part = None, 0
            if mode == RO.EF_SOLVE_DEVICE and (W.w, W.h) == (640, 480):
                try:
                    NP = 4
                    while len(trs) < NP:
                        trs.append(W.make())
                    for t_ in trs[:NP]:
                        t_.set_option(RO.EF_OPT_GRID_CTAS, sms // NP)
                        if id(t_) not in pred_of:
                            p = ModelPredictor(W.w, W.h, W.K.cx, W.K.cy, W.K.fx, W.K.fy)
                            p.stream = t_.stream
                            pred_of[id(t_)] = p
                    s_part, n_part = W.pipelined(trs[:NP], submit_sensor, args.steps, args.warmup), NP
                finally:
                    for t_ in trs:
                        t_.set_option(RO.EF_OPT_GRID_CTAS, 0)
            # tracking quality of this flow (outside the timed region): error to the ground truth over one pass
            for i in range(min(FE - 1, 40)):
                submit_sensor(trs[0], i)
                t, R = trs[0].finish()
                sensor_errs.append(float(np.linalg.norm(t - W.poses[1 + (i % (FE - 1))][:3, 3])))
            out["sensor"] = {"seconds": s_multi, "single_seconds": s_single, "partitioned_seconds": s_part, "partitions": n_part,
                             "h2d_bytes_per_step": npx * (2 + 4), "d2h_bytes_per_step": 48 + 384,
                             "handles": NE, "surfels": int(next(iter(keyframes.values())).shape[0]), "keyframe_every": KF,
                             "median_err_m": float(np.median(sensor_errs))}
        except Exception as e:  # noqa: BLE001
            out["sensor"] = {"error": repr(e)}
        for t_ in trs[1:]:
            t_.close()

    # ---- cpu baseline (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        from oracle import oracle as O
        K = W.K
        n = min(args.cpu_sample, W.F - 1)
        host = [(u16_np(W.depth[k]), W.rgba[k].cpu().numpy(), W.vmap[k].cpu().numpy(), W.nmap[k].cpu().numpy()) for k in range(n + 1)]
        cpu = O.OracleTracker(W.w, W.h, K.cx, K.cy, K.fx, K.fy, impl="cpu")
        t0 = time.perf_counter()
        for k in range(1, n + 1):
            p = W.posef[k - 1]
            cpu.init_icp_model(host[k - 1][2], host[k - 1][3], 20.0, p)
            cpu.init_rgb_model(host[k - 1][1])
            cpu.init_icp_depth(host[k][0], 20.0)
            cpu.init_rgb(host[k][1])
            cpu.get_incremental_transformation(p[:3, 3], p[:3, :3], False, W.icpw, True, False, W.so3)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"first {n} frames of the same trajectory, OpenMP C restatement (oracle/ef_oracle.c), all host cores"}
        cpu.close()
    tr.close()
    del W
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2]: 1280x720 + SO(3) pre-alignment (N=1 line only) ----
    if world == 1 and not args.no_720p and (args.width, args.height) == (640, 480):
        try:
            W2 = Workload(args, 1280, 720, 1, 101, seed, device, world)
            tr2 = W2.make()
            steps2, warm2 = min(args.steps, 200), min(args.warmup, 20)
            r2 = W2.resident(tr2, steps2, warm2, sample_clocks=False)
            ms2, calls2 = W2.solve_time(tr2, steps2, warm2)
            share2 = None if args.no_levels else W2.level_share(40)
            out["w720"] = {"ms_total": r2["ms_total"], "steps": steps2, "launches": r2["launches"], "solve_ms": ms2, "solve_calls": calls2,
                           "median_err_m": r2["median_err_m"], "level_share": share2}
            tr2.close()
            del W2
        except Exception as e:  # noqa: BLE001
            out["w720"] = {"error": repr(e)}
    return out


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world, device):
    """reference CUDA operators (oracle/_ref) under the restated RGBDOdometry host loop; inputs resident; one replica per rank."""
    from oracle import oracle as O
    so3 = bool(args.so3)
    K, poses, depth, rgba, vmap, nmap = render_sequence(args.width, args.height, args.frames, 2024 + rank, device)
    F = args.frames
    posef = poses.astype(np.float32)
    swept = None
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
        vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

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

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n, warm):
        for i in range(warm):
            step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            step(warm + i)
        barrier()
        return time.perf_counter() - t0

    clocks = ClockSampler(device)
    clocks.start()
    dt = timed(args.steps, args.warmup)
    clk = clocks.stop()
    if kind == "reference" and args.ref_sweep:
        # GPUTest-style sweep (GPUTest.cpp:247-324: every (threads, blocks) pair, all four step kernels at once, mean of five
        # calls) on a coarser grid; the best shape is then timed like the stock one
        best, best_fps = None, 0.0
        for threads in (64, 96, 128, 192, 256, 384, 512):
            for blocks in (32, 64, 96, 112, 148, 224, 296, 448):
                tr.lib.efr_tracker_set_config(tr.t, threads, blocks, threads, blocks, threads, blocks, threads, blocks)
                for i in range(2):
                    step(i)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(5):
                    step(2 + i)
                torch.cuda.synchronize()
                fps = 5 / (time.perf_counter() - t0)
                if fps > best_fps:
                    best, best_fps = (threads, blocks), fps
        tr.lib.efr_tracker_set_config(tr.t, best[0], best[1], best[0], best[1], best[0], best[1], best[0], best[1])
        swept = {"seconds": timed(args.steps, args.warmup), "threads": best[0], "blocks": best[1]}
    tr.close()
    return kind, dt, clk, swept


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

    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)

    if args.impl == "reference":
        kind, dt, clk, swept = run_reference(args, rank, world, device)
        ts = torch.tensor([dt, swept["seconds"] if swept else 0.0], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ts, op=torch.distributed.ReduceOp.MAX)
        dt, dt_swept = [float(x) for x in ts.tolist()]
        if rank == 0:
            total = args.steps * world
            fps = total / dt
            line = {"impl": "reference", "metric": "joint ICP+RGB tracking frames/s", "value": fps, "unit": "frames/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clk,
                    "reference_kind": ("reference CUDA kernels (oracle/_ref/libef_ref.so, built unmodified from the reference sources, "
                                       "GPUConfig default launch shapes) on this GPU, one replica per rank" if kind == "reference"
                                       else "OpenMP C restatement of the reference kernels on the host cores"),
                    "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count() if kind == "port" else 1, "kind": kind,
                                     "sample": f"{args.steps} frames of the same trajectory per rank"},
                    "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            if swept:
                line["value_swept"] = {"value": total / dt_swept, "unit": "frames/s", "threads": swept["threads"], "blocks": swept["blocks"],
                                       "note": "the reference at the best (threads, blocks) of a GPUTest-style sweep (GPUTest.cpp:247-324; 7 x 8 shapes, "
                                               "all four step kernels at once) instead of GPUConfig's defaults (GPUConfig.h:53-60)"}
            print(json.dumps(line))
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    r = run_ours(args, rank, world, device)

    e2e, sens = r["e2e"], r["sensor"] if r["sensor"] and "error" not in r["sensor"] else None
    ms = torch.tensor([r["ms_total"], e2e["seconds"] * 1e3 if e2e else 0.0, float(r["solve_ms"]),
                       r["concurrent"]["seconds"] * 1e3 if r["concurrent"] else 0.0, e2e["single_seconds"] * 1e3 if e2e else 0.0,
                       sens["seconds"] * 1e3 if sens else 0.0, sens["single_seconds"] * 1e3 if sens else 0.0,
                       r["batched"]["ms_total"] if r["batched"] and "error" not in r["batched"] else 0.0,
                       sens["partitioned_seconds"] * 1e3 if sens and sens.get("partitioned_seconds") else 0.0,
                       r["batched"].get("inflight2_ms_total", 0.0) if r["batched"] and "error" not in r["batched"] else 0.0], device=device, dtype=torch.float64)
    launches = torch.tensor([float(r["launches"])], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(launches, op=torch.distributed.ReduceOp.SUM)
    ms_total, e2e_ms, solve_ms, conc_ms, e2e_single_ms, sens_ms, sens_single_ms, batched_ms, sens_part_ms, batched2_ms = [float(x) for x in ms.tolist()]

    if rank == 0:
        total_frames = args.steps * world
        value = total_frames / (ms_total * 1e-3)
        avg_solve_ms = solve_ms / max(r["solve_calls"], 1)
        line = {
            "metric": "joint ICP+RGB tracking frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": r["clocks"],
            "gpu_launches": int(launches.item()),
            "us_per_gn_iteration": avg_solve_ms * 1e3 / 19.0,
            "tracking_error_m": {"median": r["median_err_m"], "max": r["max_err_m"]},
            "roofline": roofline_block(args.width, args.height, args.icp_weight, avg_solve_ms, args.solve, r["level_share"]),
        }
        if e2e:
            line["e2e"] = {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s",
                           "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                           "inflight_frames": e2e["handles"], "ctas_per_handle": e2e["ctas_per_handle"],
                           "single": {"value": total_frames / (e2e_single_ms * 1e-3), "inflight_frames": 1,
                                      "note": "one frame in flight (submit, wait for the pose, submit): the rate a closed-loop caller sees"},
                           "note": "C ABI ef_track_frame_to_model with pinned host buffers; full-GPU handles take turns: one solves while "
                                   "the others copy (PCIe bound: 12.9 MB per frame at 640x480)"}
            line["value_pipelined"] = {"value": total_frames / (conc_ms * 1e-3), "unit": "frames/s", "handles": r["concurrent"]["handles"],
                                       "ctas_per_handle": r["concurrent"]["ctas_per_handle"],
                                       "note": "inputs resident, handles on disjoint SM subsets: throughput when consecutive frames may "
                                               "overlap (`value` is the single-handle, one-frame-at-a-time rate)"}
        bt = r["batched"]
        if bt and "error" not in bt:
            fps_b = bt["frames"] * world / (batched_ms * 1e-3)
            peak, _ = measured_peaks()
            per_frame_bytes, _ = algorithmic_bytes(args.width, args.height, args.icp_weight)
            line["value_batched"] = {"value": fps_b, "unit": "frames/s", "sequences_per_launch": bt["sequences"], "ms_per_frame": 1e3 / (fps_b / world),
                                     "tracking_error_m_median": bt["median_err_m"],
                                     "two_batches_in_flight": ({"value": bt["frames"] * world / (batched2_ms * 1e-3), "unit": "frames/s"} if batched2_ms > 0 else None),
                                     "achieved_gbs_whole_frame": per_frame_bytes * fps_b / world / 1e9, "frac_whole_frame": per_frame_bytes * fps_b / world / 1e9 / peak,
                                     "note": "ef_track_frames_to_model_batch: two independent sequences per GPU, ONE persistent tracker kernel per pair of "
                                             "frames (one solver CTA per sequence; the worker CTAs alternate between the sequences, one Gauss-Newton "
                                             "iteration each, so a sequence's gather-solve-publish chain is hidden behind the other's pixels); "
                                             "inputs resident, blocking call, one pair at a time; per handle bit-identical to a single launch on the "
                                             "same number of worker CTAs. "
                                             "frac_whole_frame = algorithmic bytes of the solves / total frame time (builders and launch gaps included) / HBM peak"}
        elif bt:
            line["value_batched_error"] = bt["error"]
        if sens:
            line["e2e"]["sensor_only"] = {
                "value": total_frames / (sens_ms * 1e-3), "unit": "frames/s", "inflight_frames": sens["handles"],
                "single": {"value": total_frames / (sens_single_ms * 1e-3), "inflight_frames": 1},
                "partitioned": ({"value": total_frames / (sens_part_ms * 1e-3), "inflight_frames": sens["partitions"],
                                 "note": "the handles on disjoint SM subsets (EF_OPT_GRID_CTAS), one frame in flight each"}
                                if sens_part_ms > 0 else None),
                "h2d_bytes_per_step": sens["h2d_bytes_per_step"], "d2h_bytes_per_step": sens["d2h_bytes_per_step"],
                "surfels": sens["surfels"], "keyframe_every": sens["keyframe_every"], "tracking_error_m_median": sens["median_err_m"],
                "note": "production data flow: only depth + colour cross PCIe; the model maps are predicted on the device from a resident "
                        "key-frame surfel map (ef_op_splat_predict at the prior pose, inside the timed region) as the reference predicts "
                        "them into GL textures"}
        elif r["sensor"]:
            line["e2e_sensor_only_error"] = r["sensor"]["error"]
        if r.get("five_call"):
            line["value_reference_api"] = r["five_call"]
        w720 = r.get("w720")
        if w720 and "error" not in w720:
            ms720 = w720["solve_ms"] / max(w720["solve_calls"], 1)
            line["value_1280x720"] = {"value": w720["steps"] / (w720["ms_total"] * 1e-3), "unit": "frames/s", "ms_per_step": w720["ms_total"] / w720["steps"],
                                      "steps": w720["steps"], "gpu_launches": w720["launches"], "tracking_error_m_median": w720["median_err_m"],
                                      "workload": "BASELINE configs[2]: the same trajectory at 1280x720, SO(3) pre-alignment + joint ICP+RGB, inputs resident",
                                      "roofline": roofline_block(1280, 720, args.icp_weight, ms720, args.solve, w720["level_share"])}
        elif w720:
            line["value_1280x720_error"] = w720["error"]
        if r["cpu_baseline"]:
            line["cpu_baseline"] = r["cpu_baseline"]
        print(json.dumps(line))

    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
