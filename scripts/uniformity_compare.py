"""Chi-square uniformity of the sampler with the TRAINED fixture weights on a small satisfiable formula, side by side for
the CUDA path and the CPU oracle (reference diffusion_metrics.py:111,130-147, utils/chi_square.py).

  python scripts/uniformity_compare.py gpu    [k]    # CUDA path, fp32 and bf16 (needs a B200)
  python scripts/uniformity_compare.py oracle [k]    # CPU oracle, numpy noise (no GPU needed)

Formula: planted 3-SAT n=24 m=90 seed=2 (162 models by exact enumeration); k samples per model are drawn.  The two arms
use different noise streams (Philox on the device, numpy here), so the histograms are two independent draws from what
should be the same distribution: compare the statistics, not the counts.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from diffusionsat_b200 import synth                                    # noqa: E402
from diffusionsat_b200.graph import chains_per_reference_batch          # noqa: E402
from diffusionsat_b200.uniformity import chi_square_vs_ideal            # noqa: E402
from diffusionsat_b200.weights import load_weights                      # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "trained_small.npz")
N, M, SEED = 24, 90, 2


def report(name, hist, models, seconds, extra=""):
    chisq, p = chi_square_vs_ideal(hist, models)
    counts = np.array([hist.get(m, 0) for m in models])
    stray = sum(v for k, v in hist.items() if k not in set(models))
    print("%-12s samples %5d  models hit %3d/%d  min/median/max count %d/%d/%d  non-models %d  chi2 %.1f (dof %d) p %.3g  %.1f s %s"
          % (name, counts.sum() + stray, int((counts > 0).sum()), len(models), counts.min(), int(np.median(counts)), counts.max(),
             stray, chisq, len(models) - 1, p, seconds, extra), flush=True)


def main():
    arm = sys.argv[1] if len(sys.argv) > 1 else "gpu"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n_vars, clauses, _ = synth.planted_3sat(N, M, seed=SEED)
    models = synth.enumerate_solutions(n_vars, clauses)
    want = k * len(models)
    print("planted 3-SAT n=%d m=%d seed=%d: %d models, drawing %d SAT samples" % (N, M, SEED, len(models), want), flush=True)
    if arm == "gpu":
        from diffusionsat_b200.sampler import DiffusionSampler
        cnf = "/tmp/uniformity_compare.cnf"
        with open(cnf, "w") as fh:
            fh.write(synth.dimacs_text(n_vars, clauses))
        for precision, sampling in (("fp32", "inverse_cdf"), ("fp32_simt", "inverse_cdf"), ("bf16", "inverse_cdf"), ("fp32", "gumbel")):
            sampler = DiffusionSampler(FIXTURE, cnf, precision=precision, seed=17, sampling=sampling)
            t0 = time.time()
            hist = sampler.samples(want)
            report("cuda %s%s" % (precision, " gumbel" if sampling == "gumbel" else ""), hist, models, time.time() - t0,
                   "sat rate %.2f" % (sampler.last_stats["sat"] / sampler.last_stats["total"]))
    else:
        import torch
        from oracle import querysat_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        w = O.weights_to_torch(load_weights(FIXTURE))
        chains = chains_per_reference_batch(n_vars, len(clauses))
        rng = np.random.default_rng(17)

        def noise(_batch):
            nt = chains * n_vars
            return (torch.from_numpy(rng.random((32, nt)).astype(np.float32)), torch.from_numpy(rng.integers(0, 2, (32, nt))),
                    torch.from_numpy(rng.standard_normal((32, 32, nt, 4)).astype(np.float32)))

        t0 = time.time()
        with torch.no_grad():
            hist = O.samples(want, n_vars, clauses, w, noise, chains)
        report("oracle fp32", hist, models, time.time() - t0)


if __name__ == "__main__":
    main()
