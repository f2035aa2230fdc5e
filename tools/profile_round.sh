#!/bin/bash
# tools/profile_round.sh -- the ncu evidence of a round, on the GPU box (one GPU): outputs under gpurun_out/
#   launches_r02.csv        every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
#   launches_r02_ref.csv    the same for the reference arm (its kernels, GPUConfig default launch shapes)
#   prof_track_r02.ncu-rep  ncu --set full of k_track at 640x480 (3 launches);  prof_track720_r02.ncu-rep at 1280x720 + SO(3)
#   prof_build_r02.ncu-rep  ncu --set full of k_build_frame
# Read here with tools/profile_summarise.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --frames 12 --no-e2e --no-720p --no-levels --cpu-sample 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_[a-z]' -c 400 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/launches_r02.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel|reduceSum|pyrDown|bgr2Intensity' -c 3000 --csv --log-file gpurun_out/launches_r02_ref.csv python bench.py --impl reference --ref-sweep 0 --steps 3 --warmup 2 --frames 8 > gpurun_out/launches_r02_ref.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track -s 6 -c 3 -f -o gpurun_out/prof_track_r02 $B > gpurun_out/prof_track_r02.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track -s 6 -c 3 -f -o gpurun_out/prof_track720_r02 $B --width 1280 --height 720 --so3 1 > gpurun_out/prof_track720_r02.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_build_frame -s 6 -c 3 -f -o gpurun_out/prof_build_r02 $B > gpurun_out/prof_build_r02.out 2>&1
# second session: the alternating batched kernel (two sequences per launch) and the fused host-mode iteration at 1280x720, level 0
PYTHONPATH=. timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track_alt -s 3 -c 2 -f -o gpurun_out/prof_alt_r02 python tools/alt_run.py 2 6 > gpurun_out/prof_alt_r02.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hm_step -s 28 -c 2 -f -o gpurun_out/prof_hmstep720_r02 $B --solve host --width 1280 --height 720 > gpurun_out/prof_hmstep720_r02.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hm_residual -s 28 -c 2 -f -o gpurun_out/prof_hmres720_r02 $B --solve host --width 1280 --height 720 > gpurun_out/prof_hmres720_r02.out 2>&1
PYTHONPATH=. timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track_sym -s 8 -c 2 -f -o gpurun_out/prof_sym_r02 python tools/alt_part.py 4 1 > gpurun_out/prof_sym_r02.out 2>&1
ls -la gpurun_out/*r02*
