"""Multi-GPU execution of independent registrations: pairs shard across ranks in contiguous blocks,
a single pair never leaves its GPU, there is no collective inside a registration, and ONE all_gather of fixed-size
result records follows the batch (NCCL over NVLink on the GPU box, gloo in the CPU tests).

For pose sweeps (configs[4]: many initial poses per scan pair) the unit of sharding is the scan PAIR, so that all
inits of a pair stay on the GPU that built its clouds, covariances and label vectors once.
"""
from __future__ import annotations

import numpy as np

RECORD = 12  # doubles per registration: pose7[7], outer_iter, lm_iters_total, final_cost, n_corr_last, flags  (96 B)


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_units` owned by `rank` (sizes differ by at most one)."""
    return rank * n_units // world, (rank + 1) * n_units // world


def to_records(results) -> np.ndarray:
    out = np.zeros((len(results), RECORD), dtype=np.float64)
    for i, r in enumerate(results):
        out[i, :7] = r["pose"]
        out[i, 7:] = (r["outer_iter"], r["lm_iters_total"], r["final_cost"], r["n_corr_last"], r["flags"])
    return out


def from_records(rec: np.ndarray):
    return [dict(pose=r[:7].copy(), outer_iter=int(r[7]), lm_iters_total=int(r[8]), final_cost=float(r[9]), n_corr_last=int(r[10]),
                 flags=int(r[11])) for r in rec]


def gather_records(local: np.ndarray, units_per_rank, device=None, group=None) -> np.ndarray:
    """all_gather of the per-rank record blocks (ragged shards are padded to the largest block).

    units_per_rank[r] = number of records rank r contributes.  Returns the concatenation in rank order on every rank.
    With an uninitialised process group (single process) this is the identity."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.copy()
    world = dist.get_world_size(group)
    m = max(int(u) for u in units_per_rank)
    buf = torch.zeros((m, RECORD), dtype=torch.float64, device=device)
    if len(local):
        buf[: len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return np.concatenate([parts[r][: int(units_per_rank[r])].cpu().numpy() for r in range(world)], axis=0)


def register_sharded(register_fn, n_pairs: int, inits_per_pair: int, rank: int, world: int, device=None, group=None):
    """Run `register_fn(pair_lo, pair_hi) -> list of result dicts` (inits_per_pair results per pair, pair-major) on this
    rank's block of pairs and gather every rank's records.  Returns (all_records [n_pairs*inits_per_pair, RECORD],
    (lo, hi))."""
    lo, hi = shard_range(n_pairs, rank, world)
    local = to_records(register_fn(lo, hi)) if hi > lo else np.zeros((0, RECORD))
    assert local.shape[0] == (hi - lo) * inits_per_pair, "register_fn must return inits_per_pair results per pair"
    counts = [(shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0]) * inits_per_pair for r in range(world)]
    return gather_records(local, counts, device=device, group=group), (lo, hi)
