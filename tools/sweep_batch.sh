#!/bin/bash
# value / value_batched / value_pipelined for the product library and every build_variants/*.so
for so in instancefusion_b200/libef_track.so build_variants/*.so; do
  out=$(EF_TRACK_LIB=$PWD/$so timeout 120 python bench.py --steps 200 --warmup 20 --frames 100 --no-720p --no-levels --cpu-sample 0 "$@" 2>/dev/null | tail -1)
  echo "$so $(echo "$out" | python -c "
import sys,json
d=json.loads(sys.stdin.read())
b=d.get('value_batched') or {}
print('value', round(d['value']), 'k_track ms', round(d['roofline']['avg_launch_ms'],4), 'batched', round(b.get('value',0)), 'ms/frame', round(b.get('ms_per_frame',0),4), 'pipelined', round(d['value_pipelined']['value']), d.get('value_batched_error',''))")"
done
