"""Uniformity statistics of a sample histogram (host side, scipy).

Mirror of reference ``utils/chi_square.py:7-36`` (``chi_square_likelihood``) and of the comparison made in
``diffusion_metrics.py:111,130-147`` (histogram against the ideal ``k`` samples per solution).
"""

from __future__ import annotations

from scipy import stats


def chi_square_likelihood(observed: dict, expected: dict) -> float:
    """p-value of the chi-square test of ``observed`` against ``expected`` (dicts id -> count).
    Ids missing on one side count as zero there; a single shared id is reported as p = 1."""
    if len(observed) == 1 and len(expected) == 1:
        return 1.0
    ids = list(expected) + [k for k in observed if k not in expected]
    obs = [observed.get(k, 0) for k in ids]
    exp = [expected.get(k, 0) for k in ids]
    _, p = stats.chisquare(obs, exp)
    return p


def chi_square_vs_ideal(histogram: dict, solutions, samples_per_solution: float | None = None):
    """Chi-square statistic and p-value of a ``{solution_as_int: count}`` histogram against the uniform
    distribution over ``solutions`` (``diffusion_metrics.py:111,137``: ideal = k per solution)."""
    solutions = list(solutions)
    total = sum(histogram.get(s, 0) for s in solutions)
    if samples_per_solution is None:
        samples_per_solution = total / max(len(solutions), 1)
    obs = [histogram.get(s, 0) for s in solutions]
    exp = [samples_per_solution] * len(solutions)
    if len(solutions) < 2 or total == 0:
        return 0.0, 1.0
    chisq, p = stats.chisquare(obs, exp)
    return float(chisq), float(p)
