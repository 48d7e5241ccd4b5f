"""Multi-GPU execution of independent registrations: pairs shard across ranks, a single pair never
leaves its GPU, there is no collective inside a registration, and ONE all_gather of fixed-size result records follows
the batch (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Three ways to assign units (scan pairs) to ranks:
  * contiguous blocks (`shard_range`, the survey's default; keeps consecutive frames of a sequence together);
  * interleaved (`shard_ids(..., "interleaved")`: unit i -> rank i mod world), which spreads a slowly varying cost
    (a sequence that drives through an easy stretch) evenly — bench.py uses it;
  * dynamic chunks (`DynamicChunks`): ranks claim the next chunk from ONE shared atomic counter (the c10d store's `add`,
    a ~100 us round trip per chunk), so a rank that drew cheap pairs simply takes more — the §8(e) mitigation for
    variable pass counts; strong-scaling runs of a fixed pair list (configs[3], configs[4]) use it.

For pose sweeps (configs[4]: many initial poses per scan pair) the unit of sharding is the scan PAIR, so that all
inits of a pair stay on the GPU that built its clouds, covariances and label vectors once.
"""
from __future__ import annotations

import numpy as np

RECORD = 12  # doubles per registration: pose7[7], outer_iter, lm_iters_total, final_cost, n_corr_last, flags  (96 B)


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_units` owned by `rank` (sizes differ by at most one)."""
    return rank * n_units // world, (rank + 1) * n_units // world


def shard_ids(n_units: int, rank: int, world: int, mode: str = "block"):
    """Unit ids owned by `rank`: "block" = shard_range, "interleaved" = rank, rank + world, ..."""
    if mode == "block":
        lo, hi = shard_range(n_units, rank, world)
        return list(range(lo, hi))
    if mode == "interleaved":
        return list(range(rank, n_units, world))
    raise ValueError(mode)


class DynamicChunks:
    """Chunks [c*chunk, (c+1)*chunk) of `n_units` claimed from a shared atomic counter.

    store: a c10d store (`torch.distributed.distributed_c10d._get_default_store()` of an initialised process group);
    None = single process (a local counter).  `key` must be fresh per run.  Iterating yields (lo, hi) until the units
    are exhausted; every unit is claimed by exactly one rank."""

    def __init__(self, store, n_units: int, chunk: int, key: str = "sicp_next_chunk"):
        self.store, self.n, self.chunk, self.key = store, int(n_units), max(1, int(chunk)), key
        self._local = 0
        self.claimed = []

    def _next(self) -> int:
        if self.store is None:
            self._local += 1
            return self._local - 1
        return int(self.store.add(self.key, 1)) - 1  # add returns the new value

    def __iter__(self):
        while True:
            lo = self._next() * self.chunk
            if lo >= self.n:
                return
            hi = min(self.n, lo + self.chunk)
            self.claimed.append((lo, hi))
            yield lo, hi


def gather_by_id(local_ids, local: np.ndarray, n_units: int, per_unit: int = 1, device=None, group=None) -> np.ndarray:
    """Gather records of units owned in ANY pattern (interleaved, dynamic chunks) into global unit order.

    local_ids: unit ids this rank processed (in the order of `local`, per_unit records each).  One all_gather of the
    padded record blocks plus one of the id lists."""
    import torch
    import torch.distributed as dist

    local_ids = np.asarray(list(local_ids), dtype=np.int64)
    assert local.shape[0] == len(local_ids) * per_unit
    out = np.zeros((n_units * per_unit, RECORD), dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        for j, u in enumerate(local_ids):
            out[u * per_unit:(u + 1) * per_unit] = local[j * per_unit:(j + 1) * per_unit]
        return out
    world = dist.get_world_size(group)
    cnt = torch.tensor([len(local_ids)], dtype=torch.int64, device=device)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
    m = max(counts + [1])
    ids = torch.full((m,), -1, dtype=torch.int64, device=device)
    ids[: len(local_ids)] = torch.from_numpy(local_ids).to(ids.device)
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(all_ids, ids, group=group)
    rec = gather_records(local, [c * per_unit for c in counts], device=device, group=group)
    off = 0
    for r in range(world):
        for j in range(counts[r]):
            u = int(all_ids[r][j].item())
            out[u * per_unit:(u + 1) * per_unit] = rec[off:off + per_unit]
            off += per_unit
    return out


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
