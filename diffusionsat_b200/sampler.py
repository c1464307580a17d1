"""``DiffusionSampler(model_path, dimacs_filename).samples(n)`` on the GPU.

Mirror of reference ``satuniformity/DiffusionSampler.py``: module constants ``:17-21``,
``reverse_distribution_step_theoretic`` ``:29-37``, ``predict`` ``:40-63``, ``diffusion`` ``:78-191``,
``DiffusionSampler.__init__`` ``:197-213``, ``_prepare_checkpoints`` ``:215-227``, ``samples`` ``:229-311``.

Differences that are deliberate and documented in DESIGN.md:

* the dataset detour (solution counting with unigen/approxmc, per-copy DIMACS files, GZIP TFRecords;
  ``data/diffusion_sat_instances.py:80-94``, ``data/dimac.py:129-211``) is replaced by building the CSR/CSC
  arrays of the one formula directly; the batch composition rule is kept (``floor(max_nodes/(2n+m))``
  copies per reference batch, each its own early-exit group);
* several reference batches run in one launch; they are consumed batch by batch in order, with the
  reference's stop rules (exactly ``n`` SAT samples; abort below 0.5 % SAT rate), so the returned
  histogram does not depend on how many batches a launch holds;
* ``model_path`` is the reference's TensorFlow checkpoint directory (read without TensorFlow,
  :mod:`diffusionsat_b200.tf_checkpoint`) or a ``.npz`` from :func:`diffusionsat_b200.weights.save_weights`; a missing file prints
  "Checkpoint not found!" and continues with seeded random weights, as the reference does.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .dimacs import DimacsFile
from .graph import MAX_NODES_PER_BATCH, build_unit_graph, chains_per_reference_batch
from .query_sat import QuerySAT, distribution_at_time, randomized_rounding_tf, t_power
from .variable_assignment import VariableAssignment
from .weights import init_weights, load_weights

use_baseline_sampling = True
test_rounds = 32
diffusion_steps = 32
test_unigen = False
self_supervised = False


def reverse_distribution_step_theoretic(x, x0, t, t_increment):
    """fp32 numpy restatement of reference ``:29-37`` (``t`` is a Python float, as there)."""
    x = np.asarray(x, dtype=np.float32)
    x0 = np.asarray(x0, dtype=np.float32)
    t1 = np.power(np.float32(t), np.float32(t_power))
    t2 = np.power(np.float32(max(0.0, t - t_increment)), np.float32(t_power))
    x_new = distribution_at_time(x0, t1)
    alpha_t = (np.float32(1) - t1) / (np.float32(1) - t2)
    x_unnormed = distribution_at_time(x, np.float32(1) - alpha_t) * x_new
    return (x_unnormed / (np.sum(x_unnormed, axis=-1, keepdims=True) + np.float32(1e-8))).astype(np.float32)


def _sigmoid(z):
    z = np.asarray(z, dtype=np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-z))).astype(np.float32)


def predict(model, model_input, noisy_num, noise_scale, denoised_num=None, **noise):
    if denoised_num is not None:
        raise NotImplementedError("diffusion_step_self is unused by the sampler (self_supervised=False)")
    output = model.diffusion_step(model_input["adj_matrix"], model_input["clauses_graph"],
                                  model_input["variables_graph"], model_input.get("solutions"), noise_scale,
                                  noisy_num, **noise)
    return _sigmoid(output["prediction"]), output


def diffusion(N, model, dataset, step_data, verbose=True, prepare_image=True, *, uniforms=None, labels=None,
              normals=None):
    """Step-by-step reverse process through ``model.diffusion_step`` (one libdsat model call per step).

    ``step_data`` needs ``adjacency_matrix``, ``clauses_graph_adj``, ``variables_graph_adj``,
    ``normal_clauses`` (clause lists per graph) and ``variables_in_graph``.  ``dataset`` is unused
    (the reference only uses it to pick those keys).  Returns ``(mean cum_accuracy, predictions [N_vars],
    var_correct [N_vars])`` like the reference.  The fused whole-run kernel sequence used by
    :class:`DiffusionSampler` is ``Context.sample``; this function exists for API parity and tests.
    """
    model_input = {"adj_matrix": step_data["adjacency_matrix"], "clauses_graph": step_data["clauses_graph_adj"],
                   "variables_graph": step_data["variables_graph_adj"], "solutions": step_data.get("solutions")}
    graphs_n = [int(v) for v in step_data["variables_in_graph"]]
    n_vars = int(sum(graphs_n))
    x = np.zeros((n_vars, 2), dtype=np.float32) + np.float32(0.5)
    fixed_step = [-1] * n_vars
    fixed = [0.0] * n_vars
    cum_accuracy = np.zeros(len(graphs_n))
    predictions = None
    total_accuracy = np.zeros(len(graphs_n), dtype=bool)
    for t in range(N):
        noise_scale = 1 - t / N
        x_noisy = randomized_rounding_tf(x, noise=None if uniforms is None else uniforms[t])
        if use_baseline_sampling:
            x = x_noisy
        extra = {}
        if labels is not None:
            extra["labels"] = labels[t]
        if normals is not None:
            extra["normals"] = normals[t]
        predictions, _ = predict(model, model_input, x_noisy, noise_scale, **extra)
        x = reverse_distribution_step_theoretic(x, np.stack([1 - predictions, predictions], axis=1), noise_scale, 1 / N)
        xx = np.round(predictions)
        shift = 0
        for g, (cur_clauses, cur_n) in enumerate(zip(step_data["normal_clauses"], graphs_n)):
            asgn = VariableAssignment(n_vars=cur_n, clauses=[list(c) for c in cur_clauses])
            asgn.assign_all([bool(b) for b in xx[shift:shift + cur_n]])
            sat = asgn.satisfiable()
            total_accuracy[g] = sat
            if sat and fixed_step[shift] < 0:
                fixed[shift:shift + cur_n] = [float(v) for v in xx[shift:shift + cur_n]]
                fixed_step[shift:shift + cur_n] = [t] * cur_n
            shift += cur_n
        cum_accuracy = np.maximum(cum_accuracy, total_accuracy)
        if verbose:
            print("cum_accuracy:", np.mean(cum_accuracy), "noise_scale:", noise_scale)
    final = np.round(predictions)
    for i in range(n_vars):
        if fixed_step[i] >= 0:
            final[i] = fixed[i]
    var_correct = np.repeat(total_accuracy.astype(np.float32), graphs_n)
    return float(np.mean(cum_accuracy)), final, var_correct


def unpack_assignments(packed: np.ndarray, n_bits: int) -> list:
    """Packed little-endian 64-bit words [G, words] -> Python ints (x1 = bit 0), masked to n_bits."""
    mask = (1 << n_bits) - 1
    out = []
    for row in packed:
        value = 0
        for w, word in enumerate(row):
            value |= int(word) << (64 * w)
        out.append(value & mask)
    return out


class DiffusionSampler:
    def __init__(self, model_path, dimacs_filename, *, device: int = 0, precision: str = "fp32",
                 chains_per_launch: int | None = None, seed: int = 0, chain_offset: int = 0,
                 max_nodes_per_batch: int = MAX_NODES_PER_BATCH, verbose: bool = False, context=None):
        self.verbose = verbose
        print("model_path is ", model_path)
        weights = self._prepare_checkpoints(model_path)
        test_dimacs = DimacsFile(filename=dimacs_filename)
        test_dimacs.load()
        self.dimacs = test_dimacs
        self.n_vars = test_dimacs.number_of_vars()
        self.clauses = test_dimacs.clauses()
        self.model = QuerySAT(optimizer=None, test_rounds=test_rounds, weights=weights, device=device,
                              precision=precision, seed=seed, context=context,
                              feature_maps=weights.feature_maps, query_maps=weights.query_maps)
        self.ctx = self.model.ctx
        self.unit = build_unit_graph(self.n_vars, self.clauses)
        self.batch_chains = chains_per_reference_batch(self.n_vars, len(self.clauses), max_nodes_per_batch)
        self.chains_per_launch = chains_per_launch
        self.seed = int(seed)
        self.chain_offset = int(chain_offset)   # global id of this sampler's first chain (multi-GPU sharding)
        self.last_stats = {}

    def _prepare_checkpoints(self, model_path):
        try:
            weights = load_weights(model_path)
            print(f"Model restored from {model_path}!")
        except (FileNotFoundError, TypeError):
            print("Checkpoint not found!")
            weights = init_weights(seed=1234)
        return weights

    def _launch_chains(self, still_needed: int) -> int:
        """Chains per launch: whole reference batches, enough for the samples still needed, bounded."""
        b = self.batch_chains
        if self.chains_per_launch:
            return max(b, (self.chains_per_launch // b) * b)
        batches = max(1, -(-still_needed // b))
        rows_cap = 600_000                                  # ~12 GB of fp32 activations per launch
        cap = max(1, rows_cap // max(len(self.clauses), self.n_vars, 1) // b)
        return b * min(batches, cap)

    def samples(self, n_samples):
        """:param n_samples: how many correct samples to generate
        :return: the dict solution-as-int => count"""
        diffusion_dict = {}
        total = sat_total = 0
        still_needed = int(n_samples)
        max_lit = max((abs(l) for c in self.clauses for l in c), default=0)
        launched = 0
        stop = False
        while still_needed > 0 and not stop:
            chains = self._launch_chains(still_needed)
            if self.ctx.graph is not self.unit or self.ctx.chains != chains:
                self.ctx.set_graph(self.unit, chains=chains, group_graphs=self.batch_chains)
                self.model._graph_key = None
            packed, is_sat, _, _ = self.ctx.sample(diffusion_steps, test_rounds, seed=self.seed,
                                                   chain_offset=self.chain_offset + launched)
            launched += chains
            values = unpack_assignments(packed, self.n_vars)
            for b0 in range(0, chains, self.batch_chains):            # one reference batch at a time (:243-307)
                if still_needed == 0:
                    break
                if total > 0 and sat_total / total < 0.005:           # :261-263
                    print("too many unsat samples; stopping diffusion")
                    stop = True
                    break
                for i in range(b0, min(b0 + self.batch_chains, chains)):
                    if self.n_vars > max_lit:
                        # VariableAssignment(clauses=...) sizes its vector by the largest literal (:34-36);
                        # the reference then fails in assign_all_from_bit_list
                        raise IndexError("list assignment index out of range")
                    total += 1
                    if is_sat[i]:
                        sat_total += 1
                        key = values[i]
                        diffusion_dict[key] = diffusion_dict.get(key, 0) + 1
                        still_needed -= 1
                        if still_needed == 0:
                            break
        self.last_stats = {"total": total, "sat": sat_total, "chains_launched": launched}
        print("success rate: ", sat_total / total if total else 0.0)
        return diffusion_dict
