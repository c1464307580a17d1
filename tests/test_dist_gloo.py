"""world_size-2 gloo test of the sharding rule and the histogram merge (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionsat_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_samples(chain_ids, n_bits):
    """Deterministic stand-in for sampler output keyed by GLOBAL chain id."""
    words = -(-n_bits // 64)
    packed = np.zeros((len(chain_ids), words), dtype=np.uint64)
    is_sat = np.zeros(len(chain_ids), dtype=np.uint8)
    for i, c in enumerate(chain_ids):
        rng = np.random.default_rng(1000 + int(c) % 7)          # few distinct solutions -> collisions across ranks
        packed[i] = rng.integers(0, 2**63, size=words, dtype=np.uint64)
        packed[i, -1] &= np.uint64((1 << (n_bits - 64 * (words - 1))) - 1) if n_bits % 64 else np.uint64(2**64 - 1)
        is_sat[i] = int(c) % 3 != 0
    return packed, is_sat


def _worker(rank, world, port, total, n_bits, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = D.shard_chains(total, world, rank, multiple_of=4)
    packed, is_sat = _fake_samples(range(off, off + cnt), n_bits)
    keys, counts = D.local_histogram(packed, is_sat)
    merged = D.merge_histograms(keys, counts, n_bits)
    if rank == 0:
        torch.save(merged, out_path)
    else:
        assert merged is None
    dist.destroy_process_group()


def test_shard_chains_covers_everything():
    for total, world, mult in ((4096, 8, 31), (10, 4, 3), (5, 8, 1), (65536, 8, 12)):
        blocks = [D.shard_chains(total, world, r, mult) for r in range(world)]
        assert sum(c for _, c in blocks) == total
        pos = 0
        for off, cnt in blocks:
            assert off == pos or cnt == 0
            pos += cnt
            if cnt and off + cnt < total:
                assert cnt % mult == 0


def test_histogram_merge_world2_equals_single_process(tmp_path):
    total, n_bits = 37, 100
    out = str(tmp_path / "merged.pt")
    mp.spawn(_worker, args=(2, _free_port(), total, n_bits, out), nprocs=2, join=True)
    merged = torch.load(out, weights_only=False)
    packed, is_sat = _fake_samples(range(total), n_bits)
    keys, counts = D.local_histogram(packed, is_sat)
    single = D.merge_histograms(keys, counts, n_bits)
    assert merged == single
    assert sum(merged.values()) == int(is_sat.sum())
    assert all(0 <= k < (1 << n_bits) for k in merged)


def test_local_histogram_limit_and_empty():
    packed, is_sat = _fake_samples(range(12), 70)
    keys, counts = D.local_histogram(packed, is_sat, limit=3)
    assert counts.sum() == 3
    keys, counts = D.local_histogram(packed, np.zeros(12, dtype=np.uint8))
    assert keys.shape == (0, 2) and counts.shape == (0,)
    assert D.merge_histograms(keys, counts, 70) == {}


def test_pack_batches_follows_the_node_budget_and_loses_nothing():
    from diffusionsat_b200 import synth
    rng = np.random.default_rng(2)
    formulas = [synth.random_ksat_mixed(int(rng.integers(3, 100)), int(rng.integers(5, 400)), seed=i) for i in range(300)]
    batches = D.pack_batches(formulas, 20000)
    assert [i for b in batches for i in b] == list(range(300))          # order kept, nothing dropped
    cost = lambda i: 2 * formulas[i][0] + len(formulas[i][1])
    for k, b in enumerate(batches):
        assert sum(cost(i) for i in b) <= 20000
        if k + 1 < len(batches):                                         # greedy: the next formula did not fit
            assert sum(cost(i) for i in b) + cost(batches[k + 1][0]) > 20000
    assert D.pack_batches([(5, [[1, 2]])] * 3, max_nodes=1) == [[0], [1], [2]]      # a formula always fits an empty batch


class _FakeContext:
    """Stands in for ``_lib.Context`` in the host-logic tests below: a model call returns a deterministic function of the
    bound graph, the noisy input and the seed, so that a wrong batch -> rank assignment, a wrong gather layout or a graph
    prepared for another batch shows up as a different result."""

    def __init__(self):
        self.graph, self.chains, self.calls = None, 0, 0

    def set_graph(self, unit, chains, group_graphs=0):
        self.graph, self.chains = unit, chains

    def model_call(self, noise_scale, noisy, rounds=32, seed=0, **_):
        g = self.graph
        assert noisy.shape == (g.n_vars, 2)
        deg = np.diff(g.lit_rowptr).astype(np.float32)
        pred = deg[0::2] - deg[1::2] + noisy[:, 0] * np.float32(noise_scale) + np.float32(seed % 97) + np.float32(g.n_graphs)
        self.calls += 1
        return pred.astype(np.float32), np.array([min(rounds, 1 + g.n_clauses % 5)], dtype=np.int32), None


def _mixed_formulas(count=40):
    from diffusionsat_b200 import synth
    rng = np.random.default_rng(5)
    return [synth.random_ksat_mixed(int(rng.integers(3, 60)), int(rng.integers(5, 200)), seed=50 + i) for i in range(count)]


def _formula_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeContext()
    out = D.forward_formulas_sharded(lambda r: ctx, _mixed_formulas(), 0.4, rounds=6, seed=3, max_nodes=800)
    if rank == 0:
        torch.save((out, ctx.calls), out_path)
    else:
        assert out is None
    dist.destroy_process_group()


def test_formula_sharding_world2_equals_single_process(tmp_path):
    """Batches dealt to two ranks (each preparing its next batch on the helper thread) give the logits and step counts of
    one process; flattened formulas (graph.FlatFormula) give the same as clause lists."""
    from diffusionsat_b200 import graph as G
    formulas = _mixed_formulas()
    ctx = _FakeContext()
    logits1, steps1 = D.forward_formulas_sharded(lambda r: ctx, formulas, 0.4, rounds=6, seed=3, max_nodes=800)
    n_batches = len(D.pack_batches(formulas, 800))
    assert n_batches >= 4 and ctx.calls == n_batches
    assert [len(z) for z in logits1] == [f[0] for f in formulas]
    flat = [G.flatten_formula(*f) for f in formulas]
    logits_f, steps_f = D.forward_formulas_sharded(lambda r: _FakeContext(), flat, 0.4, rounds=6, seed=3, max_nodes=800)
    assert all(np.array_equal(a, b) for a, b in zip(logits1, logits_f)) and np.array_equal(steps1, steps_f)
    out = str(tmp_path / "formulas.pt")
    mp.spawn(_formula_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (logits2, steps2), calls_rank0 = torch.load(out, weights_only=False)
    assert calls_rank0 == -(-n_batches // 2)                            # rank 0 ran every other batch
    assert np.array_equal(steps1, steps2)
    assert all(np.array_equal(a, b) for a, b in zip(logits1, logits2))


def test_formula_sharding_propagates_a_bad_formula():
    """An out-of-range literal is found on the helper thread; the caller must see the error, not a hang or a skipped batch."""
    import pytest
    formulas = _mixed_formulas(6) + [(3, [[1, 2], [5]])]
    with pytest.raises(ValueError, match="out of range"):
        D.forward_formulas_sharded(lambda r: _FakeContext(), formulas, 0.4, rounds=2, seed=0, max_nodes=600)
