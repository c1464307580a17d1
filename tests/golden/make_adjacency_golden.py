"""Generate tests/golden/adjacency_golden.npz: the batch tensors the REFERENCE builds for a disjoint union of formulas --
``BatchedDimacsDataset.shift_clause`` + ``compute_adj_indices`` (data/dimac.py:14-18,165-170,239-241) followed by
``SatSpecifics.create_adj_matrices`` (data/SatSpecifics.py:21-69: negative literal rows offset by the batch-total variable
count, positives then negatives, graph-membership matrices) -- executed over oracle/tf_shim.py.  Pins
``UnitGraph.reference_coo``, ``unit_graph_from_reference_coo`` and the membership adapters of diffusionsat_b200.

Run in the build container only:  python tests/golden/make_adjacency_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_model_golden as M  # noqa: E402  (installs the TF stand-in, the config stub and the reference path)

mod = types.ModuleType("data.dataset")
mod.Dataset = object
sys.modules["data.dataset"] = mod
import tensorflow as tf  # noqa: E402  (the stand-in)
tf.data.Dataset = object                    # only named in type annotations of data/dimac.py
sys.modules["config"].Config.data_dir = "/tmp"
sys.modules["config"].Config.force_data_gen = False
import data.dimac as ref_dimac  # noqa: E402  (the reference's modules, unmodified)
from data.SatSpecifics import SatSpecifics  # noqa: E402

from diffusionsat_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "adjacency_golden.npz")


def main():
    cls = ref_dimac.BatchedDimacsDataset
    shifter = types.SimpleNamespace(shift_variable=cls.shift_variable)
    spec = SatSpecifics.__new__(SatSpecifics)
    out = {}
    cases = {
        "mixed": [synth.random_ksat_mixed(n, m, seed=80 + i) for i, (n, m) in enumerate([(5, 9), (12, 30), (3, 4), (8, 11)])],
        "edge": [(3, [[1, 1, -2], [], [-3]]), (2, [[-1], [1, 2]]), (4, [[4]])],        # repeats, an empty clause, unused variables
        "copies": [(6, [[1, -2, 3], [-4, 5, -6], [2, 4, 6]])] * 3,
    }
    for tag, formulas in cases.items():
        batched, offset = [], 0
        for n_vars, clauses in formulas:
            batched.extend(cls.shift_clause(shifter, [list(c) for c in clauses], offset))
            offset += n_vars
        pos, neg = ref_dimac.compute_adj_indices(batched)
        data = {
            "adj_indices_pos": torch.tensor(pos, dtype=torch.int64).reshape(-1, 2),
            "adj_indices_neg": torch.tensor(neg, dtype=torch.int64).reshape(-1, 2),
            "variable_count": torch.tensor([n for n, _ in formulas], dtype=torch.int32),
            "clauses_in_formula": torch.tensor([len(c) for _, c in formulas], dtype=torch.int32),
            "batched_clauses": torch.zeros(0), "clauses": None, "solutions": None,
        }
        res = spec.create_adj_matrices(data)
        adj, cg, vg = res["adjacency_matrix"], res["clauses_graph_adj"], res["variables_graph_adj"]
        out.update({
            tag + "_formulas": np.array([str(formulas)]),
            tag + "_adj_indices": adj.indices.numpy(), tag + "_adj_shape": np.array([int(v) for v in adj.dense_shape]),
            tag + "_cg_indices": cg.indices.numpy(), tag + "_cg_shape": np.array([int(v) for v in cg.dense_shape]),
            tag + "_vg_indices": vg.indices.numpy(), tag + "_vg_shape": np.array([int(v) for v in vg.dense_shape]),
        })
        print(tag, "adjacency", tuple(out[tag + "_adj_shape"]), "edges", len(out[tag + "_adj_indices"]))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
