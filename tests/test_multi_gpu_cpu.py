"""Host-side logic of the N>1 path on CPU (gloo, world size 2): streams shard across ranks with no
data-path collective; the only communication is the barrier and the MAX reduction of
the per-rank time that bench.py performs."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    streams = bench.shard_streams(8, rank, world)          # stream s -> rank s mod world
    t = torch.tensor([1.0 + rank])                         # pretend per-rank elapsed seconds
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    agg = bench.aggregate_throughput(steps=10, world=world, max_seconds=float(t.item()))
    if rank == 0:
        print(json.dumps({"streams0": streams, "max_t": float(t.item()), "agg": agg}))
    dist.destroy_process_group()
''') % ROOT


def test_stream_sharding_and_max_time_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["streams0"] == [0, 2, 4, 6]
    assert res["max_t"] == 2.0
    assert abs(res["agg"] - 2 * 10 / 2.0) < 1e-9


def test_reference_arm_other_ranks_exit_cleanly():
    import bench
    assert bench.shard_streams(5, 1, 2) == [1, 3]
    assert bench.aggregate_throughput(100, 4, 0.5) == 800.0
