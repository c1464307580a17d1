"""Generate tests/golden/union_golden.npz: the REFERENCE's own ``QuerySAT.diffusion_step`` (model/query_sat.py:467-481 ->
call -> loop, executed over oracle/tf_shim.py like make_model_golden.py does) on DISJOINT UNIONS OF DIFFERENT FORMULAS --
the batch shape of training and of ``predict_step`` (data/dimac.py:213-293): graphs of unequal size in one batch, so
PairNorm's per-graph statistics, the per-graph logit-map choice and the whole-batch early exit see unequal segments.

Run in the build container only:  python tests/golden/make_union_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_model_golden as M  # noqa: E402  (installs the TF stand-in and imports the reference's modules)

from diffusionsat_b200 import graph as G, synth, weights as W  # noqa: E402
from oracle import tf_shim  # noqa: E402

OUT = os.path.join(HERE, "union_golden.npz")


def union_inputs(formulas):
    union = G.build_union_graph(formulas)
    coo, shape = union.reference_coo(1)
    adj = tf_shim.SparseTensor(coo, np.ones(len(coo), np.float32), shape)
    vg_ids = np.repeat(np.arange(len(formulas)), [n for n, _ in formulas])
    cg_ids = np.repeat(np.arange(len(formulas)), [len(c) for _, c in formulas])
    vg = tf_shim.SparseTensor(np.stack([vg_ids, np.arange(union.n_vars)], 1), np.ones(union.n_vars, np.float32),
                              [len(formulas), union.n_vars])
    cg = tf_shim.SparseTensor(np.stack([cg_ids, np.arange(union.n_clauses)], 1), np.ones(union.n_clauses, np.float32),
                              [len(formulas), union.n_clauses])
    return union, adj, cg, vg


def main():
    rng = np.random.default_rng(909)
    out = {}
    cases = [
        ("u", [synth.random_ksat_mixed(n, m, seed=40 + i) for i, (n, m) in enumerate([(5, 9), (12, 30), (3, 4), (20, 70)])], 5, 0.5, 31),
        ("v", [(3, [[1, 2], [-1, 3], [2, 3]]), (2, [[1], [-1, 2]]), (4, [[1, -2, 3], [2, 2, -4], [4], [-1, -3, 4, 2]])], 6, 0.3, 32),
    ]
    for tag, formulas, rounds, noise_scale, wseed in cases:
        wts = W.init_weights(seed=wseed, bias_scale=0.1)
        model = M.build_model(wts, rounds)
        union, adj, cg, vg = union_inputs(formulas)
        n = union.n_vars
        labels = rng.integers(0, 2, n).astype(np.int32)
        normals = rng.standard_normal((rounds, n, 4)).astype(np.float32)
        uniform = rng.random(n).astype(np.float32)
        tf_shim.NOISE.clear()
        tf_shim.NOISE.uniforms.append(torch.from_numpy(uniform).reshape(n, 1))
        noisy = M.ref_sampler.randomized_rounding_tf(torch.full((n, 2), 0.5))
        tf_shim.NOISE.labels.append(torch.from_numpy(labels.astype(np.int64)))
        tf_shim.NOISE.normals.extend(torch.from_numpy(normals[r]) for r in range(rounds))
        res = model.diffusion_step(adj, cg, vg, None, noise_scale, noisy)
        out.update({
            f"{tag}_formulas": np.array([str(formulas)]), f"{tag}_rounds": rounds, f"{tag}_noise_scale": np.float32(noise_scale),
            f"{tag}_wseed": wseed, f"{tag}_labels": labels, f"{tag}_normals": normals, f"{tag}_uniform": uniform,
            f"{tag}_noisy": noisy.detach().numpy(), f"{tag}_prediction": res["prediction"].detach().numpy(),
            f"{tag}_steps_taken": int(res["steps_taken"]), f"{tag}_loss": np.float32(float(res["loss"])),
        })
        print("union", tag, "graphs", len(formulas), "variables", n, "steps_taken", int(res["steps_taken"]), "loss", float(res["loss"]))
    out.update(predict_tries_golden())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def predict_tries_golden():
    """``predict_step`` with ``prediction_tries = 3`` (reference model/query_sat.py:424-451) on the REFERENCE's class with its
    ``call`` replaced by a stub that returns three given logit vectors: pins the bookkeeping around the calls (is_graph_sat per
    try, "newly solved" clipping, per-variable masks, zeros for graphs never solved).  Stored next to the union cases."""
    rng = np.random.default_rng(4242)
    formulas = [synth.random_ksat_mixed(n, m, seed=70 + i) for i, (n, m) in enumerate([(4, 5), (6, 9), (3, 3), (5, 14), (4, 4)])]
    formulas.append((2, [[1], [-1]]))                       # never satisfiable: its variables must end as zeros
    union, adj, cg, vg = union_inputs(formulas)
    tries = 3
    logits = (rng.standard_normal((tries, union.n_vars, 1)) * 2).astype(np.float32)
    model = M.build_model(W.init_weights(seed=1, bias_scale=0.1), 2)
    model.prediction_tries = tries
    feed = iter(logits)
    model.call = lambda *a, **k: (torch.from_numpy(next(feed)), torch.tensor(0.25), 7)
    res = model.predict_step(adj, cg, vg, None)
    pred = res["prediction"].detach().numpy()
    print("predict_step tries: solved variables", int((pred != 0).sum()), "of", union.n_vars)
    return {"p_formulas": np.array([str(formulas)]), "p_logits": logits, "p_prediction": pred.astype(np.float32),
            "p_steps_taken": int(res["steps_taken"]), "p_loss": np.float32(float(res["loss"]))}


if __name__ == "__main__":
    main()
