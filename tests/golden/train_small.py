"""Fixture generator (test infrastructure, not product code): a short CPU training of the QuerySAT weights on small
random 3-SAT, so that tests have NON-random weights whose samples actually satisfy formulas (with random-init weights
the SAT rate of the sampler is 0 and `samples()` aborts, reference satuniformity/DiffusionSampler.py:261-263).

What is trained: the oracle's restatement of `QuerySAT.call(training=True, labels=solution)` (reference
model/query_sat.py:380-391: random noise level, noisy labels from the solution, loss = mean over rounds of the
cost-weighted, sorted per-graph KL of `train_loss`, :40-53, :311-315, :366-368), differentiated by torch autograd and
optimised with Adam.  It is NOT the reference's training setup (AdaBelief, 32 rounds, its dataset and schedule): the
result is only a fixture that makes SAT-rate / uniformity comparisons between the oracle and the CUDA path meaningful.

    python tests/golden/train_small.py [steps] [out.npz]

writes `tests/golden/trained_small.npz` (fp16 storage to keep the fixture small) and prints the oracle's SAT rate on
held-out formulas.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from diffusionsat_b200 import synth, weights as W        # noqa: E402
from oracle import querysat_oracle as O                   # noqa: E402


def batch_of_formulas(rng, count, n_lo=8, n_hi=20):
    formulas, labels = [], []
    for _ in range(count):
        n = int(rng.integers(n_lo, n_hi + 1))
        m = max(3, int(round(n * rng.uniform(3.0, 4.3))))
        seed = int(rng.integers(0, 2 ** 31 - 1))
        nv, clauses, hidden = synth.planted_3sat(n, m, seed=seed)
        hidden = np.asarray(hidden).astype(np.int64)
        assert all(any((hidden[abs(l) - 1] == 1) == (l > 0) for l in c) for c in clauses)
        formulas.append((nv, clauses))
        labels.append(hidden)
    return formulas, np.concatenate(labels)


def sat_rate(w, rng, n_formulas=24, n_steps=32, rounds=32, n_lo=10, n_hi=16):
    formulas, _ = batch_of_formulas(rng, n_formulas, n_lo, n_hi)
    graph = O.OracleGraph.from_formulas(formulas)
    nt = graph.n_vars
    uniforms = torch.from_numpy(rng.random((n_steps, nt)).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, 2, (n_steps, nt)))
    normals = torch.from_numpy(rng.standard_normal((n_steps, rounds, nt, 4)).astype(np.float32))
    with torch.no_grad():
        acc, final, _ = O.diffusion(n_steps, graph, w, uniforms, labels, normals, rounds)
    flags = O.graph_sat_flags(torch.from_numpy(final), graph).numpy()
    return float(flags.mean())


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "trained_small.npz")
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    init = W.init_weights(seed=1234)
    w = {k: [(a.clone().requires_grad_(True), b.clone().requires_grad_(True)) for a, b in v]
         for k, v in O.weights_to_torch(init).items()}
    params = [t for v in w.values() for pair in v for t in pair]
    opt = torch.optim.Adam(params, lr=3e-4)
    rounds = 16
    t0 = time.time()
    for step in range(steps):
        formulas, sol = batch_of_formulas(rng, 24, 8, 30)
        graph = O.OracleGraph.from_formulas(formulas)
        nt = graph.n_vars
        ns = float(rng.uniform(0.0, 1.0))                                         # :144
        labels = torch.from_numpy(sol.astype(np.int64))
        one_hot = torch.nn.functional.one_hot(labels, 2).to(torch.float32)
        at_t = O.distribution_at_time(one_hot, ns ** 0.5)                         # construct_training_input, :76-82
        noisy = O.randomized_rounding(at_t, torch.from_numpy(rng.random(nt).astype(np.float32)))
        normals = torch.from_numpy(rng.standard_normal((rounds, nt, 4)).astype(np.float32))
        _, _, loss, _, _ = O.model_loop(graph, w, ns, noisy, labels, normals, rounds)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        if step % 50 == 0 or step == steps - 1:
            print("step %4d  loss %.4f  noise %.2f  %.0f s" % (step, float(loss), ns, time.time() - t0), flush=True)
        if step % 200 == 199 or step == steps - 1:
            wd = {k: [(a.detach(), b.detach()) for a, b in v] for k, v in w.items()}
            print("  oracle SAT rate on held-out formulas: %.2f" % sat_rate(wd, np.random.default_rng(123)), flush=True)
    layers = {}
    for mlp, pairs in w.items():
        for i, (a, b) in enumerate(pairs):
            layers["%s/%d" % (mlp, i)] = (a.detach().numpy(), b.detach().numpy())
    trained = W.QuerySATWeights(type(init.layers)((k, layers[k]) for k in init.layers), 128, 128)
    blobs = {"feature_maps": np.int64(128), "query_maps": np.int64(128)}
    for name, (kernel, bias) in trained.layers.items():
        blobs[name + "/kernel"] = kernel.astype(np.float16)
        blobs[name + "/bias"] = bias.astype(np.float16)
    np.savez_compressed(out, **blobs)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
