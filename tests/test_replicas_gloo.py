"""The path does not shard: N GPUs = N independent replicas, no data-path collective.
bench.py only uses torch.distributed for the barrier and the max-over-ranks time; this test covers that
host logic with two gloo ranks on CPU."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import numpy as np
    from instancefusion_b200 import synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # every replica renders its own sequence (seed = 2024 + rank), like bench.py
    poses = synth.trajectory(4, seed=2024 + rank)
    K = synth.Intrinsics.kinect(80, 60)
    f = synth.render(poses[1], K, seed=2024 + rank, frame_id=1)
    digest = float(f["depth"].to(torch.int32).sum()) + float(poses[1][:3, 3].sum())
    ms = torch.tensor([10.0 + 5.0 * rank, 100.0 * (rank + 1)], dtype=torch.float64)   # [resident ms, e2e ms]
    frames = torch.tensor([7.0], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    if rank == 0:
        print(json.dumps({"ms": ms.tolist(), "frames": frames.item(), "distinct": len(set(digests))}))
    dist.destroy_process_group()
""" % ROOT)


def test_two_replicas_aggregate_like_bench(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
           str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["ms"] == [15.0, 200.0]      # max over ranks
    assert r["frames"] == 14.0           # sum over ranks: value = total frames / max time
    assert r["distinct"] == 2            # independent sequences per replica
