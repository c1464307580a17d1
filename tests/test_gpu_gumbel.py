"""Gumbel-argmax rounding (north star (c); the sampler sketched in reference model/query_sat.py:15-28) against the reference's
live inverse-CDF rounding floor(x0 + U) (model/query_sat.py:55-60).  The two draw DIFFERENT samples from the SAME
distribution, so the Gumbel mode is validated statistically: (1) the rounding alone on a grid of probabilities,
(2) whole sampling runs with the trained fixture weights: SAT rate and a two-sample chi-square test of the histograms."""
import os

import numpy as np
import pytest

from diffusionsat_b200 import dist as D, graph as G, synth
from diffusionsat_b200.weights import load_weights

pytestmark = pytest.mark.gpu
FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trained_small.npz")


def test_rounding_frequencies_match_the_probabilities(ctx):
    n_vars, clauses, _ = synth.planted_3sat(20, 80, seed=1)
    chains = 10000
    ctx.set_model(load_weights(FIXTURE))
    ctx.set_precision("bf16")
    ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=20)
    rows = n_vars * chains
    grid = np.array([0.02, 0.3, 0.5, 0.77, 0.98], dtype=np.float32)
    p0 = np.tile(grid, rows // len(grid))
    x = np.stack([p0, 1 - p0], axis=1).astype(np.float32)
    out = {}
    for mode in ("inverse_cdf", "gumbel"):
        ctx.set_sampling(mode)
        ctx.debug_write("X", x)
        ctx.debug_rounding(seed=5, step=3)
        r = ctx.debug_read("X")
        assert set(np.unique(r)) <= {0.0, 1.0} and np.all(r.sum(axis=1) == 1.0)
        out[mode] = r[:, 0]
        for k, p in enumerate(grid):
            sel = r[k::len(grid), 0]
            sigma = np.sqrt(p * (1 - p) / sel.size)
            assert abs(sel.mean() - p) < 5 * sigma, "%s: class-0 frequency %.4f at p=%.2f" % (mode, sel.mean(), p)
    ctx.set_sampling("inverse_cdf")
    differ = (out["gumbel"] != out["inverse_cdf"]).mean()
    assert 0.05 < differ < 0.6            # same marginals, different samples


def test_gumbel_runs_sample_the_same_distribution(ctx):
    from scipy.stats import chi2_contingency
    n_vars, clauses, _ = synth.planted_3sat(14, 50, seed=9)
    models = synth.enumerate_solutions(n_vars, clauses)
    chains = 6000
    ctx.set_model(load_weights(FIXTURE))
    ctx.set_precision("bf16")
    ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=20)
    hists, rates = {}, {}
    for mode, seed in (("inverse_cdf", 1), ("gumbel", 2), ("inverse_cdf_again", 3)):
        ctx.set_sampling("gumbel" if mode == "gumbel" else "inverse_cdf")
        ctx.sample_enqueue(32, 32, seed=seed)
        keys, counts, n_sat = ctx.hist_reduce()
        hists[mode] = D.table_to_dict(keys, counts, n_vars)
        rates[mode] = n_sat / chains
        assert set(hists[mode]) <= set(models)
    ctx.set_sampling("inverse_cdf")
    sigma = np.sqrt(rates["inverse_cdf"] * (1 - rates["inverse_cdf"]) / chains)
    assert abs(rates["gumbel"] - rates["inverse_cdf"]) < 6 * sigma + 0.01

    def p_value(a, b):
        keys = sorted(k for k in set(a) | set(b) if a.get(k, 0) + b.get(k, 0) >= 10)
        table = np.array([[a.get(k, 0) for k in keys], [b.get(k, 0) for k in keys]])
        return chi2_contingency(table)[1]
    # yardstick: two inverse-CDF runs with different seeds against each other, then Gumbel against inverse-CDF
    p_same = p_value(hists["inverse_cdf"], hists["inverse_cdf_again"])
    p_gumbel = p_value(hists["inverse_cdf"], hists["gumbel"])
    print("two-sample chi-square p-values: inverse-CDF vs inverse-CDF %.3f, Gumbel vs inverse-CDF %.3f; SAT rates %r" % (p_same, p_gumbel, rates))
    assert p_gumbel > 1e-4
