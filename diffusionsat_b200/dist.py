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


def _records(keys: np.ndarray) -> np.ndarray:
    """[K, words] uint64 (word 0 least significant) -> [K] structured array whose field order is most significant word
    first, so that numpy's lexicographic order on it is the numeric order of the encoded integers."""
    k = np.ascontiguousarray(np.asarray(keys, dtype="<u8")[:, ::-1])
    return k.view(np.dtype([("f%d" % i, "<u8") for i in range(k.shape[1])])).reshape(-1)


def _from_records(rec: np.ndarray, words: int) -> np.ndarray:
    return np.ascontiguousarray(rec.view("<u8").reshape(-1, words)[:, ::-1])


def local_histogram(packed: np.ndarray, is_sat: np.ndarray, limit: int | None = None):
    """Host restatement of ``dsat_hist_reduce`` (which the sampler uses): unique satisfying assignments of this rank,
    ``(keys [K, words] uint64 ascending by encoded integer, counts [K] int64)``.  Only SAT samples are counted
    (reference DiffusionSampler.py:297-303); ``limit`` keeps the first ``limit`` SAT samples in chain order."""
    sel = np.flatnonzero(np.asarray(is_sat) != 0)
    if limit is not None:
        sel = sel[:limit]
    words = packed.shape[1]
    if sel.size == 0:
        return np.zeros((0, words), dtype=np.uint64), np.zeros(0, dtype=np.int64)
    rec, counts = np.unique(_records(packed[sel]), return_counts=True)
    return _from_records(rec, words).astype(np.uint64), counts.astype(np.int64)


def merge_tables(tables, words: int):
    """Sum several ``(keys, counts)`` tables into one sorted table (vectorised; used across launches and ranks)."""
    tables = [(k, c) for k, c in tables if len(c)]
    if not tables:
        return np.zeros((0, words), dtype=np.uint64), np.zeros(0, dtype=np.int64)
    rec = np.concatenate([_records(k) for k, _ in tables])
    cnt = np.concatenate([np.asarray(c, dtype=np.int64) for _, c in tables])
    uniq, inverse = np.unique(rec, return_inverse=True)
    total = np.zeros(uniq.shape[0], dtype=np.int64)
    np.add.at(total, inverse.reshape(-1), cnt)
    return _from_records(uniq, words).astype(np.uint64), total


def keys_to_ints(keys: np.ndarray, n_bits: int):
    """Packed words -> Python ints (x1 = bit 0, reference utils/VariableAssignment.py:63-69), masked to ``n_bits``."""
    keys = np.asarray(keys, dtype=np.uint64)
    if keys.shape[0] == 0:
        return []
    mask = (1 << n_bits) - 1
    acc = keys[:, 0].astype(object)
    for w in range(1, keys.shape[1]):
        acc = acc | (keys[:, w].astype(object) << (64 * w))
    return [int(v) & mask for v in acc]


def table_to_dict(keys: np.ndarray, counts: np.ndarray, n_bits: int) -> dict:
    return {k: int(c) for k, c in zip(keys_to_ints(keys, n_bits), counts) if c > 0}


def merge_histograms(keys: np.ndarray, counts: np.ndarray, n_bits: int, device=None, group=None, dst: int = 0):
    """All ranks call this with their sorted unique-key table; rank ``dst`` gets the merged ``{int: count}``, the
    others ``None``.  One all-gather of the padded key tables, one reduce(SUM) of a dense count vector."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return table_to_dict(keys, counts, n_bits)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
                        else torch.device("cpu"))
    words = keys.shape[1]
    n_local = torch.tensor([keys.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(x) for x in torch.cat(sizes).cpu().tolist()]
    kmax = max(max(sizes), 1)
    padded = torch.zeros(kmax, words, dtype=torch.int64, device=device)
    if keys.shape[0]:
        padded[:keys.shape[0]] = torch.from_numpy(np.ascontiguousarray(keys).view(np.int64)).to(device)
    gathered = torch.empty(world * kmax, words, dtype=torch.int64, device=device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(gathered, padded, group=group)
    else:
        dist.all_gather(list(gathered.view(world, kmax, words).unbind(0)), padded, group=group)
    host = gathered.cpu().numpy().view(np.uint64).reshape(world, kmax, words)
    tables = [host[r, :s] for r, s in enumerate(sizes) if s > 0]
    if tables:
        table_rec = np.unique(np.concatenate([_records(t) for t in tables]))      # identical on every rank, sorted
    else:
        table_rec = _records(np.zeros((0, words), dtype=np.uint64))
    dense = torch.zeros(max(table_rec.shape[0], 1), dtype=torch.int64, device=device)
    if keys.shape[0]:
        pos = np.searchsorted(table_rec, _records(keys))      # position of each local key in the global table
        dense.index_add_(0, torch.from_numpy(pos.astype(np.int64)).to(device), torch.from_numpy(np.asarray(counts, np.int64)).to(device))
    dist.reduce(dense, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None
    total = dense.cpu().numpy()[:table_rec.shape[0]]
    return table_to_dict(_from_records(table_rec, words), total, n_bits)


def sample_chains_sharded(make_context, unit_graph, total_chains: int, batch_chains: int, n_bits: int,
                          steps: int = 32, rounds: int = 32, seed: int = 0, group=None, chains_per_launch: int | None = None,
                          return_stats: bool = False):
    """One process per GPU: every rank samples its contiguous block of the `total_chains` global chains
    (whole reference batches) in launches of at most `chains_per_launch` chains, reduces every launch's histogram on the
    device, then the ranks' tables are merged on rank 0.  `make_context(local_rank)` returns a ready `_lib.Context`
    (model set).  Returns the merged ``{int: count}`` on rank 0, ``None`` elsewhere.
    The result does not depend on the number of ranks or on the launch size: noise is keyed by the global chain id and
    launches hold whole reference batches."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offset, count = shard_chains(total_chains, world, rank, multiple_of=batch_chains)
    ctx = make_context(rank)
    words = -(-n_bits // 64)
    tables, n_sat, launches = [], 0, 0
    if chains_per_launch and count > chains_per_launch:      # equal launches of whole reference batches, none larger than the cap
        n_launch = -(-count // chains_per_launch)
        per_launch = -(-(-(-count // n_launch)) // batch_chains) * batch_chains
    else:
        per_launch = max(count, 1)
    done = 0
    while done < count:
        now = min(per_launch, count - done)
        if ctx.graph is not unit_graph or ctx.chains != now:
            ctx.set_graph(unit_graph, chains=now, group_graphs=batch_chains)
        ctx.sample_enqueue(steps, rounds, seed=seed, chain_offset=offset + done)
        keys, counts, sat = ctx.hist_reduce()
        tables.append((keys, counts))
        n_sat += sat
        done += now
        launches += 1
    keys, counts = merge_tables(tables, words)
    merged = merge_histograms(keys, counts, n_bits, group=group)
    if return_stats:
        return merged, {"chains": count, "offset": offset, "sat": n_sat, "launches": launches}
    return merged


# --------------------------------------------------------------------------- formulas sharded over ranks
def pack_batches(formulas, max_nodes: int = 20000, drop_overflow: bool = False):
    """Reference batches of a list of ``(n_vars, clauses)`` formulas (or ``graph.FlatFormula`` items): greedy packing in the
    given order while the node total ``sum(2n + m)`` stays <= ``max_nodes`` (reference ``data/dimac.py:172-174,267-293``; the
    first formula of a batch always fits).  The reference drops the formula that overflows a batch (``:281-287``, SURVEY
    Appendix A.16); here it opens the next batch, so no formula is lost.  ``drop_overflow=True`` reproduces the reference's
    batches exactly (pinned against its own code in tests/golden/batching_golden.json).  Returns a list of lists of formula
    indices."""
    batches, cur, nodes = [], [], 0
    for i, (n_vars, clauses) in enumerate(formulas):
        cost = 2 * int(n_vars) + len(clauses)
        if cur and nodes + cost > max_nodes:
            batches.append(cur)
            cur, nodes = [], 0
            if drop_overflow:
                continue
        cur.append(i)
        nodes += cost
    if cur:
        batches.append(cur)
    return batches


def forward_formulas_sharded(make_context, formulas, noise_scale: float, rounds: int = 32, seed: int = 0,
                             max_nodes: int = 20000, group=None, dst: int = 0):
    """Training-shape forward of a set of mixed formulas (BASELINE configs[3]): the formulas are packed into reference
    batches, batch b goes to rank ``b % world``, every batch is ONE disjoint-union graph and one model call (reference
    ``QuerySAT.call``, ``model/query_sat.py:133-184`` on the batches of ``data/dimac.py:213-293``), and the per-variable
    logits are gathered on rank ``dst`` (outputs are disjoint: no reduction).  Noise and the noisy inputs are keyed by
    ``(seed, batch index)``, so the result does not depend on the number of ranks.
    Returns on ``dst`` ``(logits per formula [list of float32 arrays], steps_taken per batch [int array])``."""
    from concurrent.futures import ThreadPoolExecutor
    from .graph import build_union_graph
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    batches = pack_batches(formulas, max_nodes)
    sizes = [sum(int(formulas[i][0]) for i in b) for b in batches]
    ctx = make_context(rank)
    mine = {}

    def prepare(b):
        """Host side of batch b: the union graph and the noisy input (no device work)."""
        unit = build_union_graph([formulas[i] for i in batches[b]])
        bits = np.random.default_rng([seed, b]).integers(0, 2, unit.n_vars).astype(np.float32)
        return unit, np.stack([bits, 1.0 - bits], axis=1)

    # the next batch's graph is built on a helper thread while this batch's model call runs (the call blocks in libdsat
    # with the GIL released), so the host-side build leaves the critical path
    my_batches = list(range(rank, len(batches), world))
    with ThreadPoolExecutor(max_workers=1) as pool:
        ahead = pool.submit(prepare, my_batches[0]) if my_batches else None
        for k, b in enumerate(my_batches):
            unit, noisy = ahead.result()
            ahead = pool.submit(prepare, my_batches[k + 1]) if k + 1 < len(my_batches) else None
            ctx.set_graph(unit, chains=1, group_graphs=0)            # the whole batch is one early-exit group
            pred, steps, _ = ctx.model_call(noise_scale, noisy, rounds=rounds, seed=seed + 7919 * (b + 1))
            mine[b] = (pred, int(steps[0]))
    if world == 1:
        flat = {b: mine[b] for b in mine}
    else:
        backend = dist.get_backend(group)
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        per_rank = [sum(sizes[b] + 1 for b in range(r, len(batches), world)) for r in range(world)]
        width = max(max(per_rank), 1)
        buf = torch.zeros(width, dtype=torch.float32, device=device)
        pos = 0
        for b in range(rank, len(batches), world):                # [steps_taken, logits...] per batch, back to back
            pred, steps = mine[b]
            buf[pos] = float(steps)
            buf[pos + 1:pos + 1 + sizes[b]] = torch.from_numpy(pred).to(device)
            pos += 1 + sizes[b]
        gathered = torch.empty(world * width, dtype=torch.float32, device=device)
        if backend == "nccl":
            dist.all_gather_into_tensor(gathered, buf, group=group)
        else:
            dist.all_gather(list(gathered.view(world, width).unbind(0)), buf, group=group)
        if rank != dst:
            return None
        host = gathered.cpu().numpy().reshape(world, width)
        flat = {}
        for r in range(world):
            pos = 0
            for b in range(r, len(batches), world):
                flat[b] = (host[r, pos + 1:pos + 1 + sizes[b]].copy(), int(host[r, pos]))
                pos += 1 + sizes[b]
    logits = [None] * len(formulas)
    steps_taken = np.zeros(len(batches), dtype=np.int32)
    for b, idxs in enumerate(batches):
        pred, steps_taken[b] = flat[b]
        pos = 0
        for i in idxs:
            n = int(formulas[i][0])
            logits[i] = pred[pos:pos + n]
            pos += n
    return logits, steps_taken
