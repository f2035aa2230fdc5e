#!/bin/bash
# ncu durations of the stand-alone association kernels (host-solve mode), summarised per kernel and grid
for sz in "640 480" "1280 720"; do set -- $sz
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"k_icp_step|k_rgb_residual|k_rgb_step" -s 60 -c 57 --csv --log-file gpurun_out/ops_$1.csv python bench.py --solve host --width $1 --height $2 --steps 2 --warmup 2 --frames 6 --no-e2e --cpu-sample 0 > /dev/null 2>&1
  python - $1 <<'PY'
import csv,collections,sys
rows=list(csv.reader(open('gpurun_out/ops_%s.csv'%sys.argv[1])))
for i,r in enumerate(rows):
    if r and r[0]=="ID": h=i;break
hdr=rows[h]; data=rows[h+1:]
iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value"); iM=hdr.index("Metric Name"); iG=hdr.index("Grid Size"); iID=hdr.index("ID")
per=collections.defaultdict(dict)
for r in data:
    if len(r)>iV: per[(r[iID],r[iK].split('(')[0][-24:],r[iG])][r[iM]]=float(r[iV].replace(',',''))
agg=collections.defaultdict(list)
for (i,k,g),v in per.items(): agg[(k,v.get('dram__bytes_read.sum',0)//100000)].append((v['gpu__time_duration.sum'],v.get('dram__bytes_read.sum',0),g))
print("size",sys.argv[1])
for (k,_),l in sorted(agg.items(), key=lambda kv:-kv[1][0][1]):
    t=sorted(x[0] for x in l)[len(l)//2]; b=l[0][1]
    print(f"  {k:26s} grid {l[0][2]:14s} n={len(l):2d} median {t/1000:6.2f} us  dram {b/1e6:6.2f} MB  -> {b/t:7.1f} GB/s ({100*b/t/6543.7:4.1f} % of HBM peak)")
PY
done
