python -m pytest tests/test_tracker_edge_gpu.py -m gpu -q -k graph --timeout 300 2>&1 | tail -3
for g in 0 1; do python bench.py --solve host --graph $g --steps 100 --warmup 10 --frames 60 --no-e2e --no-720p --no-levels --cpu-sample 0 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph=$g', round(d['value']), d['gpu_launches'])"; done
