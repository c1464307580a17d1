"""The model widths the constructor accepts besides the default (reference QuerySAT(feature_maps, query_maps),
model/query_sat.py:86-122: hidden sizes int(1.2 F), 4 Q, int(1.6 F) ..., all derived from F and Q): one free-running model call
per width and precision against the fp64 oracle.  feature_maps = query_maps = 64 runs the fp32-accurate tensor-core kernels
like the default; 256 exceeds what they tile and runs the same fp32 arithmetic on the CUDA cores (dsat_get_precision)."""
import numpy as np
import pytest
import torch

from diffusionsat_b200 import _lib, graph as G, synth
from oracle import querysat_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1.5e-1)])
@pytest.mark.parametrize("f,q", [(64, 64), (256, 256), (64, 128), (128, 64)])
def test_model_call_at_other_widths(ctx, f, q, precision, tol):
    n_vars, chains, rounds = 24, 7, 6
    _, clauses = synth.random_3sat(n_vars, seed=11)
    wts = H.make_weights(f, q, seed=3, bias_scale=0.1)
    ctx.set_model(wts)
    try:
        ctx.set_precision(precision)
    except _lib.DsatError:
        assert precision == "fp32"
        ctx.set_precision("fp32_simt")
    if precision == "fp32":
        assert ctx.get_precision() == (_lib.F32_TC if max(f, q) <= 128 else _lib.F32)
    ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=chains)
    n_rows = n_vars * chains
    noise = H.noise_for(n_rows, rounds, 5)
    noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
    graph, out, trace = H.oracle_trace(n_vars, clauses, chains, wts, 0.6, noisy, noise, rounds, dtype=torch.float64)
    pred, steps, loss = ctx.model_call(0.6, noisy, labels=noise["labels"], normals=noise["normals"], rounds=rounds)
    gmap = ctx.debug_groups()["graph_map"]
    same = np.repeat(gmap == trace[-1]["best_graph_map"].numpy(), n_vars)
    assert same.mean() >= 0.5
    want = out[0].numpy()
    rms = float(np.sqrt(np.mean(want ** 2)))
    err = np.abs(pred[same] - want[same])
    assert (err <= tol * np.abs(want[same]) + tol * rms).all(), "F=%d Q=%d %s: worst %.3e (rms %.3e)" % (f, q, precision, err.max(), rms)
    assert steps[0] == out[1]
