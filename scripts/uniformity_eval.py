"""Uniformity evaluation of the sampler on a small formula: exact model enumeration
(stand-in for the reference's unigen/approxmc counting, utils/AllSolutions.py:44-68), `k` samples per model as in
diffusion_metrics.py:111, chi-square against the ideal uniform histogram (utils/chi_square.py).

  python scripts/uniformity_eval.py formula.cnf [weights.npz] [k]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from diffusionsat_b200.dimacs import DimacsFile
from diffusionsat_b200.sampler import DiffusionSampler
from diffusionsat_b200.synth import enumerate_solutions
from diffusionsat_b200.uniformity import chi_square_likelihood, chi_square_vs_ideal


def main():
    cnf = sys.argv[1]
    weights = sys.argv[2] if len(sys.argv) > 2 else "no-such-checkpoint"
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    df = DimacsFile(filename=cnf)
    df.load()
    models = enumerate_solutions(df.number_of_vars(), df.clauses())
    print("models:", len(models))
    sampler = DiffusionSampler(weights, cnf, precision="bf16")
    hist = sampler.samples(len(models) * k)
    chisq, p = chi_square_vs_ideal(hist, models, samples_per_solution=None)
    print("samples:", sum(hist.values()), "distinct:", len(hist), "stats:", sampler.last_stats)
    print("chi-square vs uniform: %.2f  p = %.4g" % (chisq, p))
    ideal = {m: k for m in models}
    if sum(hist.values()) == len(models) * k:
        print("reference-style likelihood (utils/chi_square.py):", chi_square_likelihood(hist, ideal))


if __name__ == "__main__":
    main()
