"""Generate tests/golden/sat_check_golden.npz by executing the REFERENCE's own ``utils/sat.py`` (``is_graph_sat`` :165-180,
``is_batch_sat`` :118-124) over oracle/tf_shim.py on a disjoint union of mixed k-SAT formulas (with an unsatisfiable formula,
an empty clause, repeated literals and logits that are exactly 0: ``tf.round`` is half-to-even, so sigmoid(0) rounds to 0).

Run in the build container only:  python tests/golden/make_sat_check_golden.py
The host restatement (diffusionsat_b200/query_sat.py:is_graph_sat) and the oracle's ``is_batch_sat`` are checked against the
stored flags in tests/test_host_golden.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sat_check_golden.npz")

from oracle import tf_shim  # noqa: E402

tf_shim.install()
sys.path.insert(0, REF)
from utils.sat import is_batch_sat, is_graph_sat  # noqa: E402  (the reference's functions, unmodified)

from diffusionsat_b200 import graph as G, synth  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    formulas = [synth.random_ksat_mixed(int(rng.integers(3, 14)), int(rng.integers(2, 24)), seed=300 + s) for s in range(7)]
    formulas.append((3, [[1], [-1]]))                       # unsatisfiable
    formulas.append((2, [[1, 2], []]))                      # an empty clause is never satisfied
    formulas.append((3, [[1, 1, -2], [3, -3]]))             # repeated literals, a tautology
    union = G.build_union_graph(formulas)
    coo, shape = union.reference_coo(1)
    cg_ids = np.repeat(np.arange(len(formulas)), [len(c) for _, c in formulas])
    adj = tf_shim.SparseTensor(coo, np.ones(len(coo), np.float32), shape)
    cg = tf_shim.SparseTensor(np.stack([cg_ids, np.arange(union.n_clauses)], 1), np.ones(union.n_clauses, np.float32),
                              [len(formulas), union.n_clauses])
    trials = 8
    logits = rng.standard_normal((trials, union.n_vars)).astype(np.float32) * 2
    logits[np.arange(trials), rng.integers(0, union.n_vars, trials)] = 0.0
    logits[1, :] = np.abs(logits[1, :])                     # all variables true
    flags = np.stack([is_graph_sat(torch.from_numpy(z).reshape(-1, 1), adj, cg).numpy().reshape(-1) for z in logits])
    # is_batch_sat of every formula on its own, with its slice of the logits: the whole-batch flag the early exit uses
    batch = np.zeros((trials, len(formulas)), np.float32)
    off = 0
    for g, (n_vars, clauses) in enumerate(formulas):
        unit = G.build_unit_graph(n_vars, clauses)
        ucoo, ushape = unit.reference_coo(1)
        uadj = _transposed(tf_shim.SparseTensor(ucoo, np.ones(len(ucoo), np.float32), ushape))
        for t in range(trials):
            batch[t, g] = float(is_batch_sat(torch.from_numpy(logits[t, off:off + n_vars].copy()).reshape(-1, 1), uadj))
        off += n_vars
    np.savez_compressed(OUT, formulas=np.array([str(formulas)]), logits=logits, graph_sat=flags.astype(np.float32),
                        batch_sat_per_formula=batch)
    print("wrote", OUT, "graph_sat rows:", flags.astype(int).tolist())
    print("is_batch_sat per formula equals is_graph_sat:", bool(np.array_equal(batch, flags)))


def _transposed(adj):
    """``is_batch_sat`` is called with the clause x literal matrix (model/query_sat.py:188 ``cl_adj_matrix``)."""
    import tensorflow as tf
    return tf.sparse.transpose(adj)


if __name__ == "__main__":
    main()
