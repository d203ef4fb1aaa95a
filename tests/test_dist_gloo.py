"""world_size-2 gloo test of the multi-rank plumbing bench.py uses (rabe_b200/dist.py): sharding,
per-rank seeds, barrier, max-over-ranks timing and whole-job throughput."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys, torch
    sys.path.insert(0, %r)
    from rabe_b200 import dist as rd
    rank, world = rd.init("gloo")
    dev = torch.device("cpu")
    lo, hi = rd.shard(4099, rank, world)
    rd.barrier(dev)
    # rank 1 is "slower": the job time is the max, the item count the sum
    val, tmax = rd.throughput(4096, 10, 20.0 + 5.0 * rank, dev)
    mx = rd.reduce_max([float(rank), 7.0 - rank], dev)
    sm = rd.reduce_sum([float(hi - lo)], dev)
    # setup-time key broadcast (rank 0 draws, everyone receives) and the gather of fixed-size per-item outputs
    keys = rd.broadcast_bytes(b"pk-bytes-of-rank-0" * 3 if rank == 0 else b"", dev)
    g = rd.gather_fixed(torch.full((3, 2), rank, dtype=torch.uint8), dev)
    print(json.dumps({"rank": rank, "world": world, "lo": lo, "hi": hi, "val": val, "tmax": tmax, "mx": mx, "sum": sm,
                      "seed": rd.rank_seed(2, rank), "keys": keys.decode(), "gather": g.flatten().tolist()}))
    rd.finalize()
""") % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e[-2000:]
        outs.append(__import__("json").loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert [(d["lo"], d["hi"]) for d in outs] == [(0, 2050), (2050, 4099)]
    for d in outs:
        assert d["world"] == 2 and d["tmax"] == 25.0 and d["mx"] == [1.0, 7.0] and d["sum"] == [4099.0]
        assert abs(d["val"] - 2 * 4096 * 10 / 0.025) < 1e-6
    assert outs[0]["seed"] != outs[1]["seed"]
    for d in outs:
        assert d["keys"] == "pk-bytes-of-rank-0" * 3 and d["gather"] == [0] * 6 + [1] * 6


def test_shard_edges():
    from rabe_b200.dist import shard
    assert [shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard(0, 0, 1) == (0, 0)
