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


SHARDED_WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path[:0] = [%r, %r]
    from rvc_b200.sharded import ShardedIndex, shard_rows
    from oracle import knn
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(3)
    N, C, Q, k = 3001, 64, 5, 4
    index = (rng.standard_normal((N, C)) * 0.34).astype(np.float32)
    index[1700] = index[12]                                   # an exact duplicate across the shard boundary: tie -> lowest row
    queries = (rng.standard_normal((world, Q, C)) * 0.34).astype(np.float32)
    queries[1, 0] = index[12] + 1e-3
    lo, hi = shard_rows(N, rank, world)
    local = index[lo:hi]
    def search_fn(q, kk):                                      # CPU restatement of the engine's exact local top-k
        d, i = [], []
        for row in q:
            dist2 = knn.l2_f32_ordered(row, local)
            order = np.lexsort((np.arange(local.shape[0]), dist2))[:kk]
            d.append(dist2[order]); i.append(order)
        return np.stack(d), np.stack(i)
    sh = ShardedIndex(search_fn, lo, k)
    d2, idx = sh.search(queries[rank])
    want_d, want_i = [], []
    for row in queries[rank]:
        dist2 = knn.l2_f32_ordered(row, index)
        order = np.lexsort((np.arange(N), dist2))[:k]
        want_d.append(dist2[order]); want_i.append(order)
    ok = bool(np.array_equal(idx, np.stack(want_i)) and np.array_equal(d2, np.stack(want_d)))
    flags = [None] * world
    dist.all_gather_object(flags, (ok, [lo, hi], idx[0].tolist()))
    if rank == 0:
        print(json.dumps({"ok": [f[0] for f in flags], "ranges": [f[1] for f in flags], "r1q0": flags[1][2]}))
    dist.destroy_process_group()
''') % (ROOT, os.path.join(ROOT, "obs-rvc_b200"))


def test_index_sharded_knn_equals_unsharded(tmp_path):
    """Index-sharded retrieval (rvc_b200/sharded.py): all-gather queries -> local exact top-k on N / W rows -> all-gather
    candidates -> merge by (d2, row).  World size 2 on gloo with a CPU local searcher: indices AND distances equal the
    unsharded search, including a tie between duplicate rows that live on different ranks."""
    script = tmp_path / "sharded_worker.py"
    script.write_text(SHARDED_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] == [True, True], res
    assert res["ranges"] == [[0, 1501], [1501, 3001]]
    assert res["r1q0"][:2] == [12, 1700]                      # the duplicate pair: lower global row first
