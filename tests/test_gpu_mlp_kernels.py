"""Every whole-MLP kernel on its own (dsat_debug_mlp: input rows + packed weights -> output) against an fp64
restatement of reference model/mlp.py:42-50 (Dense = x @ kernel + bias, hidden activation leaky_relu 0.2), element-wise.

* fp32-accurate tensor-core path (DSAT_F32_TC, the default) and the CUDA-core fp32 path: |got - want| <= 2e-4 |want| +
  1e-4 rms(want row) per element (the reference's own fp32 arithmetic sits at ~1e-6 of the row's magnitude, the
  split-bf16 products at ~1e-5; the absolute term is per ROW because the rows are scaled differently).
* bf16 path: the fp64 chain rounds to bf16 exactly where the kernel does (operands, hidden activations after the bias add,
  leaky relu on bf16 values, outputs); 2e-3 |want| + 2e-3 rms for at least 99 % of the elements (half a bf16 ulp), and
  one bf16 ulp (8e-3) for every element: an accumulator within rounding distance of a bf16 tie may round the other way.

Sizes: n = 100, m = 428 (the bench formula's shape) with 700 chains: 547 variable tiles and 2341 clause tiles, so every
persistent CTA (or CTA pair) loops over several tiles, rings wrap, and the last tile is partial.  The reference values
are computed for rows drawn from the first, middle and last tiles and a random subset."""
import os
import subprocess
import sys

import numpy as np
import pytest

from diffusionsat_b200 import _lib, graph as G, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

F = Q = 128
MLPS = {   # name -> (input buffer, input columns (None = all), output buffer, output columns)
    "variables_query": ("VROW", F + 9, "QS", 3 * Q),
    "lit_query": ("VROW", F + 9, "LIT", 2 * Q),
    "clause_update": ("CROW", F + 2 * Q, "COUT", Q + F),
    "update_gate": ("VROW", None, "UOUT", F),
    "variables_output": ("SPRE", F, "LOGITS", 8),
}


def sample_rows(rows, rng):
    picks = [np.arange(0, min(rows, 160)), np.arange(max(rows // 2 - 100, 0), min(rows // 2 + 100, rows)),
             np.arange(max(rows - 200, 0), rows), rng.integers(0, rows, 400)]
    return np.unique(np.concatenate(picks))


def lrelu(x):
    return np.maximum(x, 0.2 * x)


def softplus64(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def reference_fp64(name, x, wts):
    """fp64 chain of one MLP on the reference's column order; x is in the kernels' packed row layout."""
    layers = wts.mlp(name)
    if name in ("variables_query", "lit_query"):
        h = x[:, :F + 9]
    elif name == "update_gate":      # packed [variables F | aux16 | grad Q | loss+ Q | loss- Q] -> reference [grad | v1 | loss+ | loss-]
        h = np.concatenate([x[:, F + 16:F + 16 + Q], x[:, :F + 9], x[:, F + 16 + Q:]], axis=1)
    else:
        h = x
    h = h.astype(np.float64)
    for i, (w, b) in enumerate(layers):
        h = h @ w.astype(np.float64) + b.astype(np.float64)
        if i + 1 < len(layers):
            h = lrelu(h)
    if name == "variables_query":
        h = np.concatenate([h, softplus64(h), softplus64(-h)], axis=1)
    return h


def reference_bf16_chain(name, x, wts):
    """The same chain with bf16 rounding at the kernel's rounding points (dsat_mlp_fused.cuh)."""
    r = H.bf16_round
    layers = wts.mlp(name)
    if name in ("variables_query", "lit_query"):
        h = x[:, :F + 9]
    elif name == "update_gate":
        h = np.concatenate([x[:, F + 16:F + 16 + Q], x[:, :F + 9], x[:, F + 16 + Q:]], axis=1)
    else:
        h = x
    h = r(h).astype(np.float64)
    for i, (w, b) in enumerate(layers):
        acc = h @ r(w).astype(np.float64) + b.astype(np.float64)
        if i + 1 < len(layers):
            hb = r(acc).astype(np.float64)                     # bias add in fp32, one rounding to bf16
            h = np.maximum(hb, r(hb * float(r(np.float32(0.2))))).astype(np.float64)    # leaky relu on bf16 values
        else:
            h = acc
    if name == "variables_query":
        h = np.concatenate([h, softplus64(h), softplus64(-h)], axis=1)
    if name != "variables_output":                             # logits stay fp32, every other output is stored as bf16
        h = r(h).astype(np.float64)
    return h


def run_case(ctx, precision, chains=700, seed=0):
    n_vars = 100
    _, clauses = synth.random_3sat(n_vars, seed=3)
    wts = H.make_weights(seed=31 + seed, bias_scale=0.1)
    ctx.set_model(wts)
    ctx.set_precision(precision)
    ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=0)
    rng = np.random.default_rng(seed)
    checked = {}
    for name, (src, _, dst, out_cols) in MLPS.items():
        rows, ld = ctx.debug_dims(src)
        x = rng.standard_normal((rows, ld)).astype(np.float32)
        if src == "VROW":
            x[:, F + 9:F + 16] = 0.0                           # aux padding columns are zero by construction
        x *= rng.uniform(0.2, 2.0, size=(rows, 1)).astype(np.float32)      # rows of different magnitude
        ctx.debug_write(src, x)
        ctx.debug_mlp(name)
        got = ctx.debug_read(dst)[:, :out_cols].astype(np.float64)
        pick = sample_rows(rows, rng)
        checked[name] = (got[pick], x[pick], wts)
    return checked


@pytest.mark.parametrize("precision", ["fp32", "fp32_simt"])
def test_fp32_mlp_kernels_elementwise(ctx, precision):
    for name, (got, x, wts) in run_case(ctx, _lib.PRECISIONS[precision]).items():
        want = reference_fp64(name, x, wts)
        rms = float(np.sqrt(np.mean(want ** 2)))
        bound = 2e-4 * np.abs(want) + 1e-4 * np.sqrt(np.mean(want ** 2, axis=1, keepdims=True))
        bad = np.abs(got - want) > bound
        assert not bad.any(), "%s (%s): %d of %d elements off, worst %.3e at value %.3e (rms %.3e)" % (
            name, precision, int(bad.sum()), bad.size, float(np.abs(got - want).max()), float(want[np.unravel_index(np.argmax(np.abs(got - want)), want.shape)]), rms)


def test_bf16_mlp_kernels_elementwise(ctx):
    for name, (got, x, wts) in run_case(ctx, _lib.BF16).items():
        want = reference_bf16_chain(name, x, wts)
        rms = float(np.sqrt(np.mean(want ** 2)))
        err = np.abs(got - want)
        tight = err <= 2e-3 * np.abs(want) + 2e-3 * rms
        loose = err <= 8e-3 * np.abs(want) + 8e-3 * rms
        assert tight.mean() >= 0.99, "%s: only %.4f of the elements within half a bf16 ulp" % (name, tight.mean())
        assert loose.all(), "%s: %d elements beyond one bf16 ulp, worst %.3e (rms %.3e)" % (name, int((~loose).sum()), float(err.max()), rms)


PLANS = {
    "x3_no_pair": {"DSAT_X3_PAIR": "0"},
    "x3_all_pair": {"DSAT_X3_PAIR": "127"},
    "x3_shallow_rings": {"DSAT_X3_A_SLOTS": "2", "DSAT_X3_W_SLOTS": "2"},
    "x3_four_epilogue_warps": {"DSAT_X3_EPI4": "127"},
    "x3_whole_clause_mlp": {"DSAT_X3_SPLIT": "0"},
    "x3_layer_per_launch_update": {"DSAT_X3_SPLIT": "3"},
    "bf16_split_off": {"DSAT_SPLIT_MODE": "0"},
    "bf16_cta_pair": {"DSAT_PAIR_MODE": "31", "DSAT_SPLIT_MODE": "0"},
    "bf16_no_pair": {"DSAT_PAIR_MODE": "0"},
    "bf16_one_tile_at_a_time": {"DSAT_PING_PONG": "0", "DSAT_A_RING": "0", "DSAT_PAIR_MODE": "0"},
}


@pytest.mark.parametrize("plan", sorted(PLANS))
def test_mlp_kernels_under_other_plans(plan):
    """The kernel plans are picked by environment switches read once per process: re-run the element-wise tests of this
    file in a fresh process under every non-default plan."""
    env = dict(os.environ)
    env.update(PLANS[plan])
    pick = "test_fp32_mlp_kernels_elementwise and fp32-" if plan.startswith("x3") else "test_bf16_mlp_kernels_elementwise"
    if plan.startswith("x3"):
        pick = "test_fp32_mlp_kernels_elementwise and not simt"
    cmd = [sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k", pick, "-p", "no:cacheprovider"]
    proc = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0, "plan %s (%s):\n%s" % (plan, PLANS[plan], proc.stdout[-3000:])
    assert "1 passed" in proc.stdout
