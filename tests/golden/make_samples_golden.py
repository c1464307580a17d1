"""Generate tests/golden/samples_golden.npz: the REFERENCE's own ``DiffusionSampler.samples`` loop
(satuniformity/DiffusionSampler.py:229-311 -- batch by batch, int encoding and SAT check through the reference's
VariableAssignment, exactly n satisfying samples, abort below 0.5 % SAT rate) executed over oracle/tf_shim.py with
``diffusion()`` replaced by a stub that returns given assignments per batch.  It pins the stop rules and the histogram that
diffusionsat_b200/sampler.py (``consume_batches`` + histogram of the consumed chains) restates in vectorised form.

Run in the build container only:  python tests/golden/make_samples_golden.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_model_golden as M  # noqa: E402  (installs the TF stand-in and imports the reference's modules)

from diffusionsat_b200 import synth  # noqa: E402
from oracle import tf_shim  # noqa: E402

OUT = os.path.join(HERE, "samples_golden.npz")
ref = M.ref_sampler


def run_reference(n_vars, clauses, batches, n_samples):
    """batches: [K, G * n_vars] 0/1 assignments, one row per reference batch of G graphs."""
    graphs = batches.shape[1] // n_vars
    step = {"variables_in_graph": [n_vars] * graphs, "clauses": [list(c) for _ in range(graphs) for c in clauses],
            "variables_graph_adj": tf_shim.SparseTensor(np.zeros((0, 2), np.int64), np.zeros(0, np.float32),
                                                        [graphs, graphs * n_vars])}
    feed = iter(batches)
    calls = {"n": 0}

    def fake_diffusion(*a, **k):
        calls["n"] += 1
        return 0.0, next(feed).astype(np.float32), None
    saved = ref.diffusion
    ref.diffusion = fake_diffusion
    fake_self = types.SimpleNamespace(data=[step] * len(batches), model=None, dataset=None)
    try:
        with contextlib.redirect_stdout(io.StringIO()) as log:
            hist = ref.DiffusionSampler.samples(fake_self, n_samples)
    finally:
        ref.diffusion = saved
    return hist, calls["n"], "too many unsat samples" in log.getvalue()


def main():
    rng = np.random.default_rng(555)
    n_vars, graphs, k_batches = 7, 6, 12
    _, clauses, _ = synth.planted_3sat(n_vars, 16, seed=9)
    out = {"n_vars": n_vars, "graphs": graphs, "clauses": np.array([str([list(c) for c in clauses])])}
    scenarios = {
        "mixed": rng.integers(0, 2, (k_batches, graphs * n_vars)),
    }
    # a first batch without any satisfying sample: the 0.5 % rule aborts before the second batch
    unsat_first = scenarios["mixed"].copy()
    bad = np.array([0] * n_vars)
    for cand in range(1 << n_vars):
        bits = np.array([(cand >> i) & 1 for i in range(n_vars)])
        if not all(any((bits[abs(l) - 1] == 1) if l > 0 else (bits[abs(l) - 1] == 0) for l in c) for c in clauses):
            bad = bits
            break
    unsat_first[0] = np.tile(bad, graphs)
    scenarios["unsat_first"] = unsat_first
    for name, batches in scenarios.items():
        out[name + "_batches"] = batches.astype(np.uint8)
        for n_samples in (1, 5, 17, 1000):
            hist, calls, aborted = run_reference(n_vars, clauses, batches, n_samples)
            keys = np.array(sorted(hist), dtype=np.int64)
            out["%s_%d_keys" % (name, n_samples)] = keys
            out["%s_%d_counts" % (name, n_samples)] = np.array([hist[int(k)] for k in keys], dtype=np.int64)
            out["%s_%d_calls" % (name, n_samples)] = calls
            out["%s_%d_aborted" % (name, n_samples)] = aborted
            print(name, "n =", n_samples, "->", sum(hist.values()), "samples,", len(hist), "distinct,", calls, "diffusion calls, aborted", aborted)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
