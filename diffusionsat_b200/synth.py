"""Seeded synthetic CNF formulas with the reference's input distributions.

* :func:`random_3sat` — cnfgen ``RandomKCNF(3, n, m)`` semantics used by reference
  ``data/CNFGen.py:39-58``: m distinct clauses, each over 3 distinct variables drawn uniformly
  without replacement, each literal negated with probability 1/2; ``m = int(4.258 n + 58.26 n^(-2/3))``
  (``data/CNFGen.py:42-43``) unless given.
* :func:`random_ksat_mixed` — clause widths of reference ``data/k_sat.py:45-46,91-93``:
  ``k = (1 w.p. 0.3 else 2) + Geometric(0.4)`` distinct variables, sign 1/2.

The reference filters for satisfiable instances with an external solver (absent here); callers that
need a SAT instance use :func:`planted_3sat`, which keeps only clauses satisfied by a hidden
assignment (the throughput benchmarks do not depend on satisfiability).
"""

from __future__ import annotations

import numpy as np


def threshold_clause_count(n_vars: int) -> int:
    return int(4.258 * n_vars + 58.26 * np.power(float(n_vars), -2.0 / 3.0))


def random_3sat(n_vars: int, n_clauses: int | None = None, seed: int = 0, k: int = 3):
    rng = np.random.default_rng(seed)
    if n_clauses is None:
        n_clauses = threshold_clause_count(n_vars)
    seen, clauses = set(), []
    while len(clauses) < n_clauses:
        vs = np.sort(rng.choice(n_vars, size=k, replace=False)) + 1
        signs = rng.integers(0, 2, size=k)
        clause = tuple(int(v if s else -v) for v, s in zip(vs, signs))
        if clause in seen:
            continue
        seen.add(clause)
        clauses.append(list(clause))
    return n_vars, clauses


def planted_3sat(n_vars: int, n_clauses: int | None = None, seed: int = 0, k: int = 3):
    """Random k-SAT conditioned on a hidden assignment being a model (always satisfiable)."""
    rng = np.random.default_rng(seed)
    if n_clauses is None:
        n_clauses = threshold_clause_count(n_vars)
    hidden = rng.integers(0, 2, size=n_vars).astype(bool)
    seen, clauses = set(), []
    while len(clauses) < n_clauses:
        vs = np.sort(rng.choice(n_vars, size=k, replace=False)) + 1
        signs = rng.integers(0, 2, size=k).astype(bool)
        if not np.any(signs == hidden[vs - 1]):
            continue
        clause = tuple(int(v if s else -v) for v, s in zip(vs, signs))
        if clause in seen:
            continue
        seen.add(clause)
        clauses.append(list(clause))
    return n_vars, clauses, hidden


def random_ksat_mixed(n_vars: int, n_clauses: int, seed: int = 0, p_k_2: float = 0.3, p_geo: float = 0.4):
    rng = np.random.default_rng(seed)
    clauses = []
    for _ in range(n_clauses):
        k = (1 if rng.random() < p_k_2 else 2) + int(rng.geometric(p_geo))
        vs = rng.choice(n_vars, size=min(n_vars, k), replace=False) + 1
        clauses.append([int(v) if rng.random() < 0.5 else -int(v) for v in vs])
    return n_vars, clauses


def dimacs_text(n_vars: int, clauses) -> str:
    lines = ["p cnf %d %d" % (n_vars, len(clauses))]
    lines += [" ".join(str(l) for l in c) + " 0" for c in clauses]
    return "\n".join(lines) + "\n"


def enumerate_solutions(n_vars: int, clauses, limit_vars: int = 24):
    """All models of a small formula as ints in the reference encoding (x1 = bit 0).

    Stand-in for reference ``utils/AllSolutions.py:44-68`` (which needs unigen/approxmc)."""
    if n_vars > limit_vars:
        raise ValueError("exact enumeration is for small formulas only")
    codes = np.arange(1 << n_vars, dtype=np.int64)
    ok = np.ones(codes.shape, dtype=bool)
    for clause in clauses:
        sat = np.zeros(codes.shape, dtype=bool)
        for lit in clause:
            bit = (codes >> (abs(lit) - 1)) & 1
            sat |= (bit == 1) if lit > 0 else (bit == 0)
        ok &= sat
    return [int(c) for c in codes[ok]]
