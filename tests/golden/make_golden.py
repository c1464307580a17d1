"""Generate tests/golden/host_golden.json by running the REFERENCE's own pure-Python modules.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference modules that are importable without TensorFlow are used as they are
(utils/DimacsFile.py, utils/VariableAssignment.py, utils/chi_square.py); data/dimac.py imports
tensorflow and pysat at module level, so those two names are stubbed with empty modules to reach its
pure-Python function compute_adj_indices (data/dimac.py:14-18).  Nothing here is shipped.
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.json")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "utils"))
    from utils.DimacsFile import DimacsFile
    from utils.VariableAssignment import VariableAssignment
    from utils.chi_square import chi_square_likelihood

    class _Anything:
        def __getattr__(self, item):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    tf = _stub("tensorflow")
    tf.__getattr__ = lambda item: _Anything()
    tf.function = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    _stub("pysat")
    _stub("pysat.solvers", Glucose4=object)
    _stub("data.dataset", Dataset=object)
    from data.dimac import compute_adj_indices
    return DimacsFile, VariableAssignment, chi_square_likelihood, compute_adj_indices


def random_formula(rng, n, m, kmax=4, dup=False):
    clauses = []
    for _ in range(m):
        k = int(rng.integers(1, kmax + 1))
        vs = rng.choice(n, size=min(k, n), replace=dup) + 1
        clauses.append([int(v) if rng.random() < 0.5 else -int(v) for v in vs])
    return clauses


def main():
    DimacsFile, VariableAssignment, chi_square_likelihood, compute_adj_indices = import_reference()
    rng = np.random.default_rng(2024)
    gold = {}

    texts = [
        "c comment\np cnf 3 2\n1 -2 0\n2 3 0\n",
        "p cnf 2 1\n1 2 0 5 6\n",                       # tokens after the first 0 are ignored
        "p cnf 2 2\n1 -2 0\n0\n",                        # a bare 0 adds an empty clause
        "p cnf 2 1\n1 7 0\n",                            # literal larger than the header raises n_vars
        "x p cnf 4 1\n1 2 0\n",                          # 'p cnf' matched anywhere; first 5 chars dropped
        "p cnf 3 1\n-- odd comment\nv 1 -2 3\n1 2 3 0\n",
        "p cnf 3 2\n1 2 0\n%\n0\n",                      # SATLIB trailer -> ValueError
        "\n\n  p cnf 5 3 \n 1  -5 0\n\n-3 0\n 2 4 -1 0\n",
        "p cnf 3\n1 0\n",                                # no blank after the count: last char is dropped
    ]
    cases = []
    for text in texts:
        df = DimacsFile()
        try:
            df.load_from_string(text)
            cases.append({"text": text, "n_vars": df.number_of_vars(), "clauses": df.clauses(),
                          "b_values": {str(k): v for k, v in df.b_values.items()}, "str": str(df)})
        except Exception as exc:  # noqa: BLE001 - we record the type
            cases.append({"text": text, "error": type(exc).__name__})
    gold["dimacs_parse"] = cases

    red = [[[1, 2], [1, 2], [3]], [[1, 2, 3], [1, 2], [3]], [[1, 2, -3], [1, -2], [1]]]   # utils/test_DimacsFile.py
    for _ in range(20):
        red.append(random_formula(rng, int(rng.integers(3, 8)), int(rng.integers(2, 14))))
    out = []
    for clauses in red:
        df = DimacsFile(clauses=[list(c) for c in clauses])
        df.reduce_clauses()
        out.append({"clauses": clauses, "reduced_sorted": sorted(df.clauses(), key=lambda c: (len(c), c)),
                    "lengths": [len(c) for c in df.clauses()]})
    gold["reduce_clauses"] = out

    va = []
    for _ in range(30):
        n = int(rng.integers(1, 140))
        clauses = random_formula(rng, n, int(rng.integers(1, 30)), dup=bool(rng.integers(0, 2)))
        clauses.append([n])                      # make the largest literal n so the vector has n entries
        bits = [int(b) for b in rng.integers(0, 2, size=n)]
        a = VariableAssignment(clauses=clauses)
        a.assign_all_from_bit_list(bits)
        b = VariableAssignment(n, [])
        b.assign_all_from_int(int(a))
        va.append({"n": n, "clauses": clauses, "bits": bits, "int": str(int(a)), "sat": bool(a.satisfiable()),
                   "str": str(a), "int_list": a.as_int_list(), "roundtrip": b.values() == a.values()})
    a = VariableAssignment(3, [])
    a.assign_all_from_int_list([1, 2, 3])        # utils/VariableAssignment.py:109-112 -> 7
    va.append({"n": 3, "clauses": [], "bits": [1, 1, 1], "int": str(int(a)), "sat": True, "str": str(a),
               "int_list": a.as_int_list(), "roundtrip": True})
    gold["variable_assignment"] = va

    adj = []
    for _ in range(12):
        n = int(rng.integers(2, 30))
        clauses = random_formula(rng, n, int(rng.integers(1, 40)), dup=True)
        pos, neg = compute_adj_indices(clauses)
        adj.append({"n": n, "clauses": clauses, "pos": pos, "neg": neg})
    adj.append({"n": 3, "clauses": [[1, 1, -2], [], [-3]], "pos": compute_adj_indices([[1, 1, -2], [], [-3]])[0],
                "neg": compute_adj_indices([[1, 1, -2], [], [-3]])[1]})
    gold["adj_indices"] = adj

    chi = []
    pairs = [({123: 1, 124: 1, 125: 2}, {123: 1, 124: 1, 125: 1, 126: 1}), ({1: 5}, {1: 5})]
    for _ in range(6):
        k = int(rng.integers(2, 9))
        exp = {i: 10 for i in range(k)}
        obs_counts = rng.multinomial(10 * k, np.ones(k) / k)
        pairs.append(({i: int(c) for i, c in enumerate(obs_counts)}, exp))
    for obs, exp in pairs:
        with contextlib.redirect_stdout(io.StringIO()):
            p = chi_square_likelihood(dict(obs), dict(exp))
        chi.append({"observed": {str(k): v for k, v in obs.items()}, "expected": {str(k): v for k, v in exp.items()},
                    "p": float(p)})
    gold["chi_square"] = chi

    with open(OUT, "w") as handle:
        json.dump(gold, handle, indent=0, sort_keys=True)
    print("wrote", OUT, {k: len(v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
