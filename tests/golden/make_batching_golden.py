"""Generate tests/golden/batching_golden.json by running the REFERENCE's own batching code (data/dimac.py:
``BatchedDimacsDataset.sat_node_count`` :172-174, ``__batch_files`` :267-293, ``shift_clause`` :165-170) with tensorflow /
pysat stubbed as in make_golden.py.  Pins ``graph.chains_per_reference_batch`` (copies of one formula per batch),
``dist.pack_batches`` (mixed formulas; the reference drops the formula that overflows a batch, SURVEY Appendix A.16 --
the golden records which ones) and the variable shift of ``graph.build_union_graph``.

Run in the build container only:  python tests/golden/make_batching_golden.py
"""
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

OUT = os.path.join(HERE, "batching_golden.json")


def main():
    make_golden.import_reference()                      # stubs tensorflow / pysat, puts the reference on sys.path
    cfg = types.ModuleType("config")
    cfg.Config = type("Config", (), {"data_dir": "/tmp", "force_data_gen": False, "max_nodes_per_batch": 20000,
                                     "input_mode": "literals"})
    sys.modules["config"] = cfg
    import importlib
    import data.dimac as ref_dimac
    ref_dimac = importlib.reload(ref_dimac)
    cls = ref_dimac.BatchedDimacsDataset
    fake = types.SimpleNamespace(max_nodes_per_batch=20000)
    fake.shift_variable = cls.shift_variable
    batch_files = cls._BatchedDimacsDataset__batch_files
    random.shuffle = lambda x: None                     # keep the reference's batches in creation order

    gold = {"copies": [], "mixed": [], "shift": []}
    rng = np.random.default_rng(11)
    # copies of one formula (what the sampler's dataset produces, data/diffusion_sat_instances.py:80-94)
    for n_vars, n_clauses in [(30, 133), (100, 428), (250, 1065), (3, 2), (9000, 3000), (10000, 43000), (12, 19976)]:
        nodes = cls.sat_node_count(fake, n_vars, n_clauses)
        for max_nodes in (20000, 5000):
            fake.max_nodes_per_batch = max_nodes
            files = [(nodes, "f%d" % i) for i in range(3 * max(1, max_nodes // nodes) + 5)]
            batches = batch_files(fake, files)
            gold["copies"].append({"n_vars": n_vars, "n_clauses": n_clauses, "max_nodes": max_nodes, "nodes": nodes,
                                   "first_batch": len(batches[0]), "batch_sizes": [len(b) for b in batches[:3]]})
    # mixed formulas
    for trial in range(6):
        sizes = [(int(rng.integers(3, 101)), int(rng.integers(1, 450))) for _ in range(int(rng.integers(20, 400)))]
        max_nodes = int(rng.choice([20000, 3000, 700]))
        fake.max_nodes_per_batch = max_nodes
        files = [(cls.sat_node_count(fake, n, m), i) for i, (n, m) in enumerate(sizes)]
        batches = batch_files(fake, files)
        kept = [i for b in batches for i in b]
        gold["mixed"].append({"sizes": sizes, "max_nodes": max_nodes, "batches": batches,
                              "dropped": sorted(set(range(len(sizes))) - set(kept))})
    # variable shift of a formula placed at an offset
    for off in (0, 7, 250):
        clauses = [[1, -2, 3], [-1], [2, 2, -3]]
        gold["shift"].append({"offset": off, "clauses": clauses, "shifted": cls.shift_clause(fake, clauses, off)})
    json.dump(gold, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "copies per batch:",
          [(c["n_vars"], c["max_nodes"], c["first_batch"]) for c in gold["copies"]])
    print("dropped by the reference:", [len(m["dropped"]) for m in gold["mixed"]], "of", [len(m["sizes"]) for m in gold["mixed"]])


if __name__ == "__main__":
    main()
