"""Worker of tests/test_shard.py: one rank of a world_size-N gloo job exercising the host-side sharding + gather
logic of semantic-icp_b200/python/shard.py with a deterministic stand-in for the per-pair registration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fake_result(pair, j):
    rng = np.random.default_rng(1000 * pair + j)
    q = rng.normal(size=4)
    return dict(pose=np.concatenate([q / np.linalg.norm(q), rng.normal(size=3)]), outer_iter=3 + pair % 5, lm_iters_total=40 + j, final_cost=float(pair) + 0.25 * j,
                n_corr_last=1000 + pair, flags=pair & 1)


def main():
    import torch.distributed as dist
    import semantic_icp_b200 as pkg

    n_pairs, inits = int(sys.argv[1]), int(sys.argv[2])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    seen = []

    def register_fn(lo, hi):
        seen.append((lo, hi))
        return [fake_result(p, j) for p in range(lo, hi) for j in range(inits)]

    rec, (lo, hi) = pkg.shard.register_sharded(register_fn, n_pairs, inits, rank, world)
    exp = pkg.shard.to_records([fake_result(p, j) for p in range(n_pairs) for j in range(inits)])
    assert rec.shape == exp.shape, (rec.shape, exp.shape)
    assert np.array_equal(rec, exp), "gathered records differ from the single-process result"
    assert (lo, hi) == pkg.shard.shard_range(n_pairs, rank, world)
    back = pkg.shard.from_records(rec)
    assert back[-1]["outer_iter"] == 3 + (n_pairs - 1) % 5
    # interleaved assignment: unit i -> rank i mod world, gathered back into global pair order
    ids = pkg.shard.shard_ids(n_pairs, rank, world, "interleaved")
    local = pkg.shard.to_records([fake_result(p, j) for p in ids for j in range(inits)]) if ids else np.zeros((0, pkg.shard.RECORD))
    rec2 = pkg.shard.gather_by_id(ids, local, n_pairs, per_unit=inits)
    assert np.array_equal(rec2, exp), "interleaved gather differs from the single-process result"
    # dynamic chunks: one shared atomic counter in the c10d store; every pair is claimed exactly once
    from torch.distributed.distributed_c10d import _get_default_store

    dc = pkg.shard.DynamicChunks(_get_default_store(), n_pairs, chunk=2, key="test_chunks")
    mine = [p for lo2, hi2 in dc for p in range(lo2, hi2)]
    local = pkg.shard.to_records([fake_result(p, j) for p in mine for j in range(inits)]) if mine else np.zeros((0, pkg.shard.RECORD))
    rec3 = pkg.shard.gather_by_id(mine, local, n_pairs, per_unit=inits)
    assert np.array_equal(rec3, exp), "dynamic-chunk gather differs from the single-process result"
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world} ok pairs [{lo},{hi})")


if __name__ == "__main__":
    main()
