#!/bin/bash
# run the short resident bench for the product library and every build_variants/*.so; prints fps / frame ms / k_track ms
# and checks the pose against the product library's (first line) on the same frames
for so in instancefusion_b200/libef_track.so build_variants/*.so; do
  out=$(EF_TRACK_LIB=$PWD/$so timeout 120 python bench.py --steps 300 --warmup 30 --frames 100 --no-e2e --no-720p --no-levels --cpu-sample 0 "$@" 2>/dev/null | tail -1)
  echo "$so $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), 'k_track', round(d['roofline']['avg_launch_ms'],4), 'err', round(d['tracking_error_m']['median']*1e6,1), round(d['tracking_error_m']['max']*1e6,1))")"
done
