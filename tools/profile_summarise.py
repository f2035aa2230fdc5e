#!/usr/bin/env python
"""Turn the ncu captures of tools/profile_round.sh (gpurun_out/*_r02*) into the committed summaries under profiles/:
  r02_launch_list.txt            per-kernel launches and device-time shares of one tracked frame, ours and the reference
  r02_k_track_ncu_full.txt       the counters of the persistent tracker kernel (640x480 and 1280x720 + SO(3))
  r02_k_build_frame_ncu_full.txt the same for the one-launch builder
  roofline_traffic.json          DRAM and L2 bytes of one k_track launch per image size (bench.py's roofline.traffic / l2_bytes)
Runs on the CPU box (ncu -i needs no GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_table(path, skip_warm):
    rows = list(csv.reader(open(path, errors="ignore")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    iK, iV, iM, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    seq = []
    for r in rows[h + 1:]:
        if len(r) > iV and r[iM] == "gpu__time_duration.sum":
            import re
            m = re.search(r"(k_[a-z0-9_]+|[A-Za-z0-9_]*Kernel[A-Za-z0-9_]*|reduceSum|pyrDown[A-Za-z]*|bgr2Intensity[A-Za-z]*)", r[iK])
            name = m.group(1) if m else r[iK][:40]
            seq.append((name, float(r[iV].replace(",", "")), r[iG], r[iB]))
    return seq


def summarise_launches():
    out = []
    for tag, fn, per_frame_kernel in (("ours (bench.py: ef_track_frame_to_model, device solve)", "launches_r02.csv", "k_track"),
                                      ("reference arm (oracle/_ref: the reference's kernels, GPUConfig default shapes)", "launches_r02_ref.csv", None)):
        path = os.path.join(G, fn)
        if not os.path.exists(path):
            continue
        seq = launch_table(path, 0)
        # keep only the product's / reference's kernels (torch's render kernels have other names)
        if per_frame_kernel:
            mine = [s for s in seq if s[0].startswith("k_")]
            frames = max(1, sum(1 for s in mine if s[0].startswith("k_track")))
        else:
            mine = [s for s in seq if not s[0].startswith("k_") and "at::" not in s[0]]
            frames = max(1, sum(1 for s in mine if s[0].startswith("copyMaps")))
        agg = collections.OrderedDict()
        for name, ns, g, b in mine:
            a = agg.setdefault(name, [0, 0.0, set()])
            a[0] += 1
            a[1] += ns
            a[2].add(f"{g} x {b}")
        total = sum(a[1] for a in agg.values())
        out.append(f"== {tag}: {len(mine)} launches over {frames} tracked frames = {len(mine) / frames:.1f} launches per frame, "
                   f"{total / frames / 1000:.1f} us of kernel time per frame (ncu: cold-cache, serialised -- compare shares)")
        for name, (n, ns, shapes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"  {name:34s} {n / frames:6.1f} /frame  {ns / n / 1000:8.2f} us each  {100 * ns / total:5.1f} % of kernel time   grid x block {sorted(shapes)[:3]}")
        out.append("")
    open(os.path.join(P, "r02_launch_list.txt"), "w").write(
        "ncu --metrics gpu__time_duration.sum --clock-control none (tools/profile_round.sh), 640x480 joint ICP+RGB\n\n" + "\n".join(out))
    print("\n".join(out))


METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
           "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__inst_executed_pipe_fp64.sum",
           "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum"]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def full_summary(rep_name, title, out_name, traffic_key=None, traffic=None):
    rep = os.path.join(G, rep_name)
    if not os.path.exists(rep):
        return
    recs, units = raw_page(rep)
    lines = [title, f"source: gpurun_out/{rep_name} (ncu --set full --clock-control none --import-source on, 3 launches; values of the LAST launch, "
             "the launches before it warm the caches only as far as a serialised profile allows)", ""]
    r = recs[-1]
    lines.append(f"kernel: {r.get('Kernel Name', '?')[:110]}")
    for m in METRICS:
        if m in r:
            lines.append(f"  {m:86s} {r[m]:>16s} {units.get(m, '')}")
    dram = float(r.get("dram__bytes_read.sum", "0").replace(",", "")) + float(r.get("dram__bytes_write.sum", "0").replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    dram_b = (float(r.get("dram__bytes_read.sum", "0").replace(",", "")) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
              + float(r.get("dram__bytes_write.sum", "0").replace(",", "")) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0))
    lts_b = float(r.get("lts__t_sectors.sum", "0").replace(",", "")) * 32.0  # = lts__t_bytes (32-byte sectors; the full set has no byte counter)
    lines.append("")
    lines.append(f"per launch: DRAM read + write {dram_b / 1e6:.2f} MB, L2 (lts__t_sectors x 32 B = lts__t_bytes) {lts_b / 1e6:.2f} MB")
    open(os.path.join(P, out_name), "a" if os.path.exists(os.path.join(P, out_name)) and traffic_key and traffic else "w").write("\n".join(lines) + "\n\n")
    if traffic is not None and traffic_key:
        traffic[traffic_key] = {"dram_bytes": int(dram_b), "lts_bytes": int(lts_b), "source": f"profiles/{out_name}"}
    print("\n".join(lines[:4]), "...", lines[-1])


def main():
    summarise_launches()
    traffic = {}
    for f in ("r02_k_track_ncu_full.txt",):
        if os.path.exists(os.path.join(P, f)):
            os.remove(os.path.join(P, f))
    full_summary("prof_track_r02.ncu-rep", "k_track, 640x480 joint ICP+RGB (256 threads x 148 CTAs, one launch = 19 Gauss-Newton iterations)", "r02_k_track_ncu_full.txt", "640x480", traffic)
    full_summary("prof_track720_r02.ncu-rep", "k_track, 1280x720 SO(3) + joint ICP+RGB (384 threads x 148 CTAs)", "r02_k_track_ncu_full.txt", "1280x720", traffic)
    full_summary("prof_build_r02.ncu-rep", "k_build_frame, 640x480 (every pyramid of a frame-to-model frame from one launch)", "r02_k_build_frame_ncu_full.txt")
    full_summary("prof_alt_r02.ncu-rep", "k_track_alt, 640x480 joint ICP+RGB, TWO sequences per launch (2 solver CTAs + 146 workers alternating between the sequences)", "r02_k_track_alt_ncu_full.txt")
    full_summary("prof_sym_r02.ncu-rep", "k_track_sym, 640x480 joint ICP+RGB, a handle on 37 of the 148 SMs (symmetric body: every CTA gathers all rows and solves)", "r02_k_track_sym_ncu_full.txt")
    for f in ("r02_host_fused_ncu_full.txt",):
        if os.path.exists(os.path.join(P, f)):
            os.remove(os.path.join(P, f))
    hm = {"appending": True}  # (a non-empty dict makes full_summary append the second kernel to the same file)
    full_summary("prof_hmstep720_r02.ncu-rep", "k_hm_step (host-solve mode, fused iteration: icpStep + rgbStep + reduction), 1280x720 level 0", "r02_host_fused_ncu_full.txt", "hm_step_1280x720", hm)
    full_summary("prof_hmres720_r02.ncu-rep", "k_hm_residual (host-solve mode, fused iteration: computeRgbResidual -> 8-byte correspondences), 1280x720 level 0", "r02_host_fused_ncu_full.txt", "hm_res_1280x720", hm)
    if traffic:
        json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
        print(traffic)


if __name__ == "__main__":
    sys.exit(main())
