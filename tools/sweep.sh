#!/bin/bash
# run the short resident bench for the product library and every build_variants/*.so; prints fps / frame ms / k_track ms
for so in instancefusion_b200/libef_track.so build_variants/*.so; do
  out=$(EF_TRACK_LIB=$PWD/$so timeout 90 python bench.py --steps 200 --warmup 20 --frames 100 --no-e2e --cpu-sample 0 "$@" 2>/dev/null | tail -1)
  echo "$so $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['avg_launch_ms'],4))")"
done
