"""Generate tests/golden/model_golden.npz by executing the REFERENCE's own sources
(model/query_sat.py, model/mlp.py, layers/normalization.py, loss/sat.py, utils/sat.py,
metrics/sat_metrics.py, satuniformity/DiffusionSampler.py) over oracle/tf_shim.py, a torch-backed
stand-in for the TensorFlow calls they make (TensorFlow itself cannot be installed offline).

Run in the build container only:  python tests/golden/make_model_golden.py
Everything random is injected through tf_shim.NOISE; inputs and outputs are stored so that the CPU oracle
(tests/test_oracle_vs_reference_golden.py) and the CUDA path (tests/test_gpu_parity.py) can be checked
against what the reference code itself computed.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.npz")

from oracle import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, REF)

# modules of the reference that only exist to pull in unavailable third-party packages
for name, attrs in {
    "optimization": {}, "optimization.AdaBelief": {"AdaBeliefOptimizer": lambda **k: None},
    "data.diffusion_sat_instances": {"DiffusionSatDataset": object},
}.items():
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
cfg = types.ModuleType("config")
cfg.Config = type("Config", (), {"learning_rate": 2e-4, "input_mode": "literals", "max_nodes_per_batch": 20000})
sys.modules["config"] = cfg

from model.query_sat import QuerySAT  # noqa: E402  (the reference's class, unmodified)
import satuniformity.DiffusionSampler as ref_sampler  # noqa: E402
from metrics.sat_metrics import SATAccuracyTF  # noqa: E402,F401

from diffusionsat_b200 import graph as G, synth, weights as W  # noqa: E402


def make_inputs(n_vars, clauses, chains):
    unit = G.build_unit_graph(n_vars, clauses)
    coo, shape = unit.reference_coo(chains)
    adj = tf_shim.SparseTensor(coo, np.ones(len(coo), np.float32), shape)
    n, m = n_vars * chains, len(clauses) * chains
    vg_ids = np.repeat(np.arange(chains), n_vars)
    cg_ids = np.repeat(np.arange(chains), len(clauses))
    vg = tf_shim.SparseTensor(np.stack([vg_ids, np.arange(n)], 1), np.ones(n, np.float32), [chains, n])
    cg = tf_shim.SparseTensor(np.stack([cg_ids, np.arange(m)], 1), np.ones(m, np.float32), [chains, m])
    return adj, cg, vg


def build_model(wts, rounds):
    model = QuerySAT(optimizer=None, test_rounds=rounds)
    for mlp, name in ((model.variables_query, "variables_query"), (model.lit_mlp, "lit_query"),
                      (model.clause_mlp, "clause_update"), (model.update_gate, "update_gate"),
                      (model.variables_output, "variables_output")):
        layers = wts.mlp(name)
        assert len(layers) == len(mlp.dense_layers)
        for dense, (k, b) in zip(mlp.dense_layers, layers):
            dense.set_weights(k, b)
    return model


class Ragged:
    def __init__(self, flat):
        self.flat_values = flat


def main():
    rng = np.random.default_rng(77)
    out = {}
    cases = [("a", 12, None, 3, 4, 0.625, 5), ("b", 30, None, 2, 6, 0.25, 6), ("c", 3, [[1, 2], [-1, 3], [2, 3]], 4, 5, 0.4, 7),
             ("d", 5, [[1, -2, 3], [2, 2, -4], [4], [-1, -3, 4, 2]], 2, 3, 0.9, 8)]
    for tag, n_vars, clauses, chains, rounds, noise_scale, wseed in cases:
        if clauses is None:
            _, clauses = synth.random_3sat(n_vars, seed=wseed)
        wts = W.init_weights(seed=wseed, bias_scale=0.1)
        model = build_model(wts, rounds)
        adj, cg, vg = make_inputs(n_vars, clauses, chains)
        n = n_vars * chains
        labels = rng.integers(0, 2, n).astype(np.int32)
        normals = rng.standard_normal((rounds, n, 4)).astype(np.float32)
        uniform = rng.random(n).astype(np.float32)
        x_half = torch.full((n, 2), 0.5)
        tf_shim.NOISE.clear()
        tf_shim.NOISE.uniforms.append(torch.from_numpy(uniform).reshape(n, 1))
        noisy = ref_sampler.randomized_rounding_tf(x_half)            # reference model/query_sat.py:55-60
        tf_shim.NOISE.labels.append(torch.from_numpy(labels.astype(np.int64)))
        tf_shim.NOISE.normals.extend(torch.from_numpy(normals[r]) for r in range(rounds))
        res = model.diffusion_step(adj, cg, vg, None, noise_scale, noisy)   # reference :467-481 -> call -> loop
        used = rounds - len(tf_shim.NOISE.normals)
        out.update({
            f"step_{tag}_n_vars": n_vars, f"step_{tag}_clauses": np.array([str(clauses)]), f"step_{tag}_chains": chains,
            f"step_{tag}_rounds": rounds, f"step_{tag}_noise_scale": np.float32(noise_scale), f"step_{tag}_wseed": wseed,
            f"step_{tag}_labels": labels, f"step_{tag}_normals": normals, f"step_{tag}_uniform": uniform,
            f"step_{tag}_noisy": noisy.detach().numpy(), f"step_{tag}_prediction": res["prediction"].detach().numpy(),
            f"step_{tag}_steps_taken": int(res["steps_taken"]), f"step_{tag}_loss": np.float32(float(res["loss"])),
            f"step_{tag}_rounds_run": used,
        })
        print("diffusion_step", tag, "steps_taken", int(res["steps_taken"]), "loss", float(res["loss"]), "rounds run", used)

    # posterior step alone (reference DiffusionSampler.py:29-37)
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.0, 1.0]])
    p = torch.tensor([0.9, 0.2, 0.5, 0.731])
    x0 = torch.stack([1 - p, p], dim=1)
    for i, t in enumerate([1.0, 0.75, 0.5, 1 / 32]):
        out[f"post_{i}_t"] = np.float64(t)
        out[f"post_{i}_out"] = ref_sampler.reverse_distribution_step_theoretic(x, x0, t, 1 / 32).numpy()
    out["post_x"], out["post_p"] = x.numpy(), p.numpy()

    # the reference's whole diffusion() loop on a batch of copies
    for tag, n_vars, n_clauses, chains, steps, rounds, wseed in (("e", 8, 16, 4, 5, 3, 21), ("f", 14, 40, 3, 4, 3, 22)):
        _, clauses, _ = synth.planted_3sat(n_vars, n_clauses, seed=wseed)
        wts = W.init_weights(seed=wseed, bias_scale=0.1)
        ref_sampler.test_rounds = rounds
        model = build_model(wts, rounds)
        adj, cg, vg = make_inputs(n_vars, clauses, chains)
        n = n_vars * chains
        uniforms = rng.random((steps, n)).astype(np.float32)
        labels = rng.integers(0, 2, (steps, n)).astype(np.int32)
        normals = rng.standard_normal((steps, rounds, n, 4)).astype(np.float32)
        tf_shim.NOISE.clear()
        # queue order follows the reference: per step one uniform draw, one label draw, then normals per round; the
        # model may break early, so normals are queued per step through a hook on randomized rounding
        state = {"t": 0}
        orig_rr = ref_sampler.randomized_rounding_tf

        def rr(xx, _state=state):
            t = _state["t"]
            tf_shim.NOISE.normals.clear()                       # drop unused normals of an early-exited step
            tf_shim.NOISE.uniforms.append(torch.from_numpy(uniforms[t]).reshape(n, 1))
            tf_shim.NOISE.labels.append(torch.from_numpy(labels[t].astype(np.int64)))
            tf_shim.NOISE.normals.extend(torch.from_numpy(normals[t, r]) for r in range(rounds))
            _state["t"] = t + 1
            return orig_rr(xx)

        ref_sampler.randomized_rounding_tf = rr
        step_data = {
            "adjacency_matrix": adj, "clauses_graph_adj": cg, "variables_graph_adj": vg,
            "solutions": Ragged(torch.zeros(n, dtype=torch.int32)),
            "clauses": [c for _ in range(chains) for c in clauses], "normal_clauses": [clauses] * chains,
            "variables_in_graph": [n_vars] * chains,
        }
        dataset = types.SimpleNamespace(args_for_train_step=lambda sd: {
            "adj_matrix": sd["adjacency_matrix"], "clauses_graph": sd["clauses_graph_adj"],
            "variables_graph": sd["variables_graph_adj"], "solutions": sd["solutions"]})
        with contextlib.redirect_stdout(io.StringIO()):
            acc, predictions, _ = ref_sampler.diffusion(steps, model, dataset, step_data, verbose=False, prepare_image=False)
        ref_sampler.randomized_rounding_tf = orig_rr
        out.update({
            f"diff_{tag}_n_vars": n_vars, f"diff_{tag}_clauses": np.array([str(clauses)]), f"diff_{tag}_chains": chains,
            f"diff_{tag}_steps": steps, f"diff_{tag}_rounds": rounds, f"diff_{tag}_wseed": wseed,
            f"diff_{tag}_uniforms": uniforms, f"diff_{tag}_labels": labels, f"diff_{tag}_normals": normals,
            f"diff_{tag}_predictions": np.asarray(predictions, dtype=np.float32), f"diff_{tag}_accuracy": np.float64(acc),
        })
        print("diffusion", tag, "cum accuracy", acc, "predictions", np.asarray(predictions)[:n_vars])

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
