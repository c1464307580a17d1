"""Sharding of independent chains over ranks and the one collective of the path: the histogram merge.

The reference is single-process (no collective anywhere).  Chains are independent
(PairNorm, logit-map choice and SAT checks are per graph), so rank r simply owns the chains
``[offset_r, offset_r + count_r)`` — noise is keyed by the global chain id, so the union of all ranks'
samples equals a single-GPU run of the same chains.  Only the final ``{solution_as_int: count}``
histogram crosses GPUs:

1. every rank sorts/uniques its satisfying assignments (packed 64-bit words) locally,
2. ``all_gather`` of the padded unique-key tables (NCCL over NVLink on GPUs, gloo in CPU tests),
3. every rank builds the identical global sorted key table and scatters its counts into a dense vector,
4. ``reduce(SUM)`` of that vector to rank 0, which materialises the dict.

Keys travel as int64 reinterpretations of the uint64 words; torch.distributed is plumbing here.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_chains(total_chains: int, world_size: int, rank: int, multiple_of: int = 1):
    """Contiguous block of chains for ``rank``: ``(offset, count)``; blocks are multiples of
    ``multiple_of`` (whole reference batches) except possibly the last."""
    units = -(-total_chains // multiple_of)
    per = -(-units // world_size)
    lo = min(rank * per * multiple_of, total_chains)
    hi = min((rank + 1) * per * multiple_of, total_chains)
    return lo, hi - lo


def local_histogram(packed: np.ndarray, is_sat: np.ndarray, limit: int | None = None):
    """Unique satisfying assignments of this rank: ``(keys [K, words] uint64 sorted, counts [K] int64)``.
    Only SAT samples are counted (reference DiffusionSampler.py:297-303); ``limit`` keeps the first
    ``limit`` SAT samples in chain order."""
    sel = np.flatnonzero(np.asarray(is_sat) != 0)
    if limit is not None:
        sel = sel[:limit]
    words = packed.shape[1]
    if sel.size == 0:
        return np.zeros((0, words), dtype=np.uint64), np.zeros(0, dtype=np.int64)
    keys, counts = np.unique(packed[sel], axis=0, return_counts=True)
    return keys.astype(np.uint64), counts.astype(np.int64)


def keys_to_ints(keys: np.ndarray, n_bits: int):
    mask = (1 << n_bits) - 1
    out = []
    for row in keys:
        value = 0
        for w, word in enumerate(row):
            value |= int(word) << (64 * w)
        out.append(value & mask)
    return out


def merge_histograms(keys: np.ndarray, counts: np.ndarray, n_bits: int, device=None, group=None, dst: int = 0):
    """All ranks call this; rank ``dst`` gets the merged ``{int: count}``, the others ``None``."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(zip(keys_to_ints(keys, n_bits), (int(c) for c in counts)))
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
                        else torch.device("cpu"))
    words = keys.shape[1]
    n_local = torch.tensor([keys.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    kmax = max(max(sizes), 1)
    padded = torch.zeros(kmax, words, dtype=torch.int64, device=device)
    if keys.shape[0]:
        padded[:keys.shape[0]] = torch.from_numpy(keys.view(np.int64).copy()).to(device)
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    tables = [g[:s].cpu().numpy().view(np.uint64) for g, s in zip(gathered, sizes) if s > 0]
    if tables:
        table = np.unique(np.concatenate(tables, axis=0), axis=0)      # identical on every rank
    else:
        table = np.zeros((0, words), dtype=np.uint64)
    dense = torch.zeros(max(table.shape[0], 1), dtype=torch.int64, device=device)
    if keys.shape[0]:
        # position of each local key in the global table (rows are sorted lexicographically)
        lookup = {row.tobytes(): i for i, row in enumerate(table)}
        pos = torch.tensor([lookup[row.tobytes()] for row in keys], dtype=torch.int64, device=device)
        dense.index_add_(0, pos, torch.from_numpy(counts).to(device))
    dist.reduce(dense, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    total = dense.cpu().numpy()
    return {k: int(c) for k, c in zip(keys_to_ints(table, n_bits), total[:table.shape[0]]) if c > 0}


def sample_chains_sharded(make_context, unit_graph, total_chains: int, batch_chains: int, n_bits: int,
                          steps: int = 32, rounds: int = 32, seed: int = 0, group=None):
    """One process per GPU: every rank samples its contiguous block of the `total_chains` global chains
    (whole reference batches), then the histograms are merged on rank 0.  `make_context(local_rank)` returns a
    ready `_lib.Context` (model set).  Returns the merged ``{int: count}`` on rank 0, ``None`` elsewhere.
    The result does not depend on the number of ranks: noise is keyed by the global chain id."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offset, count = shard_chains(total_chains, world, rank, multiple_of=batch_chains)
    ctx = make_context(rank)
    if count > 0:
        ctx.set_graph(unit_graph, chains=count, group_graphs=batch_chains)
        packed, is_sat, _, _ = ctx.sample(steps, rounds, seed=seed, chain_offset=offset)
        keys, counts = local_histogram(packed, is_sat)
    else:
        words = -(-n_bits // 64)
        keys, counts = np.zeros((0, words), dtype=np.uint64), np.zeros(0, dtype=np.int64)
    return merge_histograms(keys, counts, n_bits, group=group)
