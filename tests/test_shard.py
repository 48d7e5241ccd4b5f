"""Multi-GPU path on CPU: world_size-2 (and 3, ragged) gloo jobs run the sharding + single all_gather logic."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_everything(pkg):
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 4, 8):
            blocks = [pkg.shard.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_and_dynamic_assignment(pkg):
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            ids = [pkg.shard.shard_ids(n, r, world, "interleaved") for r in range(world)]
            assert sorted(i for b in ids for i in b) == list(range(n))
            assert max(len(b) for b in ids) - min(len(b) for b in ids) <= 1
            assert [i for r in range(world) for i in pkg.shard.shard_ids(n, r, world, "block")] == list(range(n))
    dc = pkg.shard.DynamicChunks(None, 11, 4)  # single process: a local counter
    assert list(dc) == [(0, 4), (4, 8), (8, 11)] and dc.claimed == [(0, 4), (4, 8), (8, 11)]
    res = [dict(pose=np.arange(7.0) + i, outer_iter=i, lm_iters_total=i, final_cost=0.0, n_corr_last=0, flags=0) for i in (2, 0, 1)]
    rec = pkg.shard.gather_by_id([2, 0, 1], pkg.shard.to_records(res), 3)
    assert [int(r[7]) for r in rec] == [0, 1, 2]


def test_records_round_trip(pkg):
    res = [dict(pose=np.arange(7) + i, outer_iter=i, lm_iters_total=10 * i, final_cost=0.5 * i, n_corr_last=i * i, flags=i & 1) for i in range(5)]
    rec = pkg.shard.to_records(res)
    assert rec.shape == (5, pkg.shard.RECORD) and rec.itemsize * rec.shape[1] == 96
    back = pkg.shard.from_records(rec)
    assert all(np.array_equal(a["pose"], b["pose"]) and a["outer_iter"] == b["outer_iter"] for a, b in zip(res, back))
    assert np.array_equal(pkg.shard.gather_records(rec, [5]), rec)  # no process group: identity


@pytest.mark.parametrize("world,n_pairs,inits", [(2, 8, 1), (2, 5, 3), (3, 4, 2)])
def test_sharded_batch_gloo(world, n_pairs, inits):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), str(n_pairs), str(inits)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count(" ok pairs ") == world
