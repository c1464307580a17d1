"""Tests with the TRAINED fixture weights (tests/golden/trained_small.npz, produced by tests/golden/train_small.py).

With random-init weights the sampler never satisfies a formula (SAT rate 0, `samples()` aborts as the reference does),
so the first-SAT latch, the batch early exit, the SAT-only histogram and the uniformity statistic are exercised only on
trivial formulas.  These weights make them bite on real 3-SAT."""
import os

import numpy as np
import pytest
import torch

from diffusionsat_b200 import graph as G, synth
from diffusionsat_b200.weights import load_weights
from oracle import querysat_oracle as O
from tests import helpers as H

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "trained_small.npz")


def trained():
    return load_weights(FIXTURE)


def _oracle_run(n_vars, clauses, chains, wts, noise, steps, rounds):
    graph = O.OracleGraph.copies(n_vars, clauses, chains)
    trace = []
    acc, final, latch = O.diffusion(steps, graph, O.weights_to_torch(wts), torch.from_numpy(noise["uniforms"]),
                                    torch.from_numpy(noise["labels"].astype(np.int64)), torch.from_numpy(noise["normals"]),
                                    rounds, trace=trace)
    return graph, final, latch, trace


def test_fixture_loads_and_oracle_solves_small_formulas():
    wts = trained()
    assert (wts.feature_maps, wts.query_maps) == (128, 128) and wts.n_params() == 856788
    n_vars, clauses, _ = synth.planted_3sat(14, 50, seed=9)
    chains, steps, rounds = 6, 12, 8
    noise = H.noise_for(n_vars * chains, rounds, 3, steps=steps)
    graph, final, latch, trace = _oracle_run(n_vars, clauses, chains, wts, noise, steps, rounds)
    flags = O.graph_sat_flags(torch.from_numpy(final), graph).numpy()
    assert flags.mean() >= 0.5                       # random-init weights: 0.0
    assert (latch >= 0).any()                        # the first-SAT latch fired
    assert any(t["steps_taken"] < rounds - 1 for t in trace)     # and so did the batch early exit


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp32_simt"])
@pytest.mark.parametrize("n_vars,n_clauses,chains,group,seed", [(14, 50, 6, 6, 9), (12, 44, 6, 6, 0), (20, 80, 8, 4, 1), (16, 60, 300, 3, 2)])
def test_cuda_sample_matches_oracle_with_trained_weights(ctx, n_vars, n_clauses, chains, group, seed, precision):
    """Bit-exact reverse diffusion under injected noise where formulas DO get satisfied: first-SAT latch, per-group
    early exit (steps_taken < rounds) and SAT flags all take their non-trivial branches.  Groups that exit early are
    skipped by every later kernel of the model call (tile-wise in the whole-MLP kernels: the 300-chain case has tiles that
    are skipped, tiles that straddle finished and live groups, and live tiles), so this is also the test of that skipping."""
    from diffusionsat_b200 import _lib
    from diffusionsat_b200.sampler import unpack_assignments
    _, clauses, _ = synth.planted_3sat(n_vars, n_clauses, seed=seed)
    wts = trained()
    steps, rounds = 10, 6
    ctx.set_model(wts)
    ctx.set_precision(_lib.PRECISIONS[precision])
    ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=group)
    noise = H.noise_for(n_vars * chains, rounds, 50 + seed, steps=steps)
    packed, is_sat, latch_step, sat_any = ctx.sample(steps, rounds, uniforms=noise["uniforms"], labels=noise["labels"],
                                                     normals=noise["normals"])
    got = unpack_assignments(packed, n_vars)
    checked = satisfied = 0
    oracle_groups = range(0, chains, group) if chains <= 16 else list(range(0, chains, group))[::9]
    for g0 in oracle_groups:                         # the oracle runs one early-exit group (reference batch) at a time
        rows = slice(g0 * n_vars, (g0 + group) * n_vars)
        sub = dict(uniforms=noise["uniforms"][:, rows], labels=noise["labels"][:, rows], normals=noise["normals"][:, :, rows])
        graph, final, latch, trace = _oracle_run(n_vars, clauses, group, wts, sub, steps, rounds)
        for c in range(group):
            bits = final[c * n_vars:(c + 1) * n_vars]
            margins = [np.abs(t["predictions"].numpy()[c * n_vars:(c + 1) * n_vars] - 0.5).min() for t in trace]
            if min(margins) < 1e-3:
                continue                              # a probability on a rounding boundary: fp32 order effects decide
            checked += 1
            assert got[g0 + c] == O.encode_assignment(bits)
            assert latch_step[g0 + c] == latch[c * n_vars]
            ok = O._satisfiable_py([bool(b) for b in bits], clauses)
            assert bool(is_sat[g0 + c]) == ok
            satisfied += ok
    assert checked >= (len(oracle_groups) * group) // 2 and satisfied >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sampler_returns_only_models_and_enough_of_them(ctx, tmp_path, precision):
    """`DiffusionSampler.samples(n)` end to end on satisfiable 3-SAT: exactly n samples, every key a model."""
    from diffusionsat_b200.sampler import DiffusionSampler
    n_vars, clauses, _ = synth.planted_3sat(14, 50, seed=9)
    cnf = tmp_path / "f.cnf"
    cnf.write_text(synth.dimacs_text(n_vars, clauses))
    sampler = DiffusionSampler(FIXTURE, str(cnf), context=ctx, precision=precision, seed=11)
    hist = sampler.samples(200)
    assert sum(hist.values()) == 200
    assert sampler.last_stats["sat"] / sampler.last_stats["total"] >= 0.3
    models = set(synth.enumerate_solutions(n_vars, clauses))
    assert set(hist) <= models
    assert len(hist) >= min(3, len(models))           # it samples, it does not collapse onto one assignment
