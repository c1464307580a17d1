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
    from .dist import keys_to_ints
    return keys_to_ints(np.asarray(packed, dtype=np.uint64), n_bits)


def consume_batches(is_sat: np.ndarray, batch: int, still_needed: int, total: int, sat_total: int,
                    min_sat_rate: float = 0.005):
    """The reference's stop rules (``satuniformity/DiffusionSampler.py:243-307``) applied to the SAT flags of one launch,
    vectorised.  Chains are consumed reference batch by reference batch, in order; before each batch the run stops if
    ``sat_total / total < min_sat_rate`` (``:261-263``); inside a batch it stops right after the sample that completes
    ``still_needed`` (``:305-307``).  Returns ``(chains_consumed, sat_counted, stopped_on_rate)``: the histogram of the
    launch is that of chains ``[0, chains_consumed)``."""
    is_sat = np.asarray(is_sat) != 0
    chains = is_sat.shape[0]
    csum = np.concatenate([[0], np.cumsum(is_sat, dtype=np.int64)])
    pos = 0
    while pos < chains and still_needed > 0:
        if total > 0 and sat_total / total < min_sat_rate:
            return pos, int(csum[pos]), True
        end = min(pos + batch, chains)
        sat_here = int(csum[end] - csum[pos])
        if sat_here >= still_needed:        # the sample that completes the request ends the run inside this batch
            cut = int(np.searchsorted(csum, csum[pos] + still_needed, side="left"))     # first index with that many SAT
            return cut, int(csum[cut]), False
        total += end - pos
        sat_total += sat_here
        still_needed -= sat_here
        pos = end
    return pos, int(csum[pos]), False


class DiffusionSampler:
    """Drop-in for the reference class.  ``precision="fp32"`` (default) runs the Dense layers fp32-accurately on the
    tensor cores; ``"bf16"`` is the faster, stated-separately path.

    Reproducibility: noise is a counter-based Philox stream keyed by ``(seed, global chain id)``.  Every ``samples()``
    call continues with fresh chains (the reference draws fresh TF randomness per call too); a new sampler with the same
    ``seed`` and ``chain_offset`` replays the same chains, and ``reset_chains()`` rewinds this one."""

    def __init__(self, model_path, dimacs_filename, *, device: int = 0, precision: str = "fp32",
                 chains_per_launch: int | None = None, seed: int = 0, chain_offset: int = 0,
                 max_nodes_per_batch: int = MAX_NODES_PER_BATCH, verbose: bool = False, context=None,
                 sampling: str = "inverse_cdf"):
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
        self.ctx.set_sampling(sampling)         # "gumbel": Gumbel-argmax rounding (same distribution, other samples)
        self.unit = build_unit_graph(self.n_vars, self.clauses)
        self.batch_chains = chains_per_reference_batch(self.n_vars, len(self.clauses), max_nodes_per_batch)
        self.chains_per_launch = chains_per_launch
        self.seed = int(seed)
        self.chain_offset = int(chain_offset)   # global id of this sampler's first chain (multi-GPU sharding)
        self.min_sat_rate = 0.005               # reference :261-263; 0 disables the abort (throughput runs)
        self._chains_consumed = 0               # chains launched by earlier samples() calls
        self.last_stats = {}

    def reset_chains(self, chain_offset: int | None = None):
        """Rewind the chain counter (and optionally move the block of global chain ids this sampler draws from)."""
        self._chains_consumed = 0
        if chain_offset is not None:
            self.chain_offset = int(chain_offset)

    def _prepare_checkpoints(self, model_path):
        """Reference ``:215-227``: restore the latest checkpoint, or print "Checkpoint not found!" and go on with
        (seeded) random weights.  Only a missing path counts as "not found": a file that exists but cannot be parsed
        raises, so a reader bug can never turn into silently random weights."""
        import os
        if model_path is None or not (os.path.exists(str(model_path)) or os.path.exists(str(model_path) + ".index")):
            print("Checkpoint not found!")
            return init_weights(seed=1234)
        try:
            weights = load_weights(model_path)
        except FileNotFoundError:              # a directory without a `checkpoint` file / ckpt-N.index
            print("Checkpoint not found!")
            return init_weights(seed=1234)
        print(f"Model restored from {model_path}!")
        return weights

    def _launch_chains(self, still_needed: int, chains_left: int | None = None, sat_rate: float | None = None) -> int:
        """Chains per launch: whole reference batches, bounded by the activation memory.  The reference runs one batch at a
        time until it has enough samples; a launch here holds as many batches as the samples still needed are expected to
        take at the SAT rate seen so far (first launch: 50 %, at least four batches -- small launches are latency-bound, so
        spare chains cost nothing), and batches are consumed in order with the reference's stop rules, so the histogram does
        not depend on the launch size."""
        b = self.batch_chains
        if self.chains_per_launch:
            n = max(b, (self.chains_per_launch // b) * b)
        else:
            rate = max(sat_rate if sat_rate is not None else 0.5, 0.02)
            batches = max(4, -(-int(still_needed / rate * 1.15 + 1) // b))
            rows_cap = 900_000                              # variable/clause rows per launch (~20 GB of activations at n=100)
            cap = max(1, rows_cap // max(len(self.clauses), self.n_vars, 1) // b)
            n = b * min(batches, cap)
            if self.ctx.graph is self.unit and n <= self.ctx.chains <= 2 * n:
                n = self.ctx.chains                         # same shape as the previous launch: buffers and the captured step are kept
        if chains_left is not None:
            n = min(n, max(chains_left, 1))
            if chains_left - n < b:             # a remainder smaller than one reference batch rides along as a last, partial group
                n = max(chains_left, 1)
        return n

    def samples(self, n_samples, *, max_chains: int | None = None):
        """:param n_samples: how many correct samples to generate
        :param max_chains: (extension) stop after this many chains even if fewer samples were found
        :return: the dict solution-as-int => count"""
        from .dist import table_to_dict
        keys, counts = self.samples_table(n_samples, max_chains=max_chains)
        return table_to_dict(keys, counts, self.n_vars)

    def samples_table(self, n_samples, *, max_chains: int | None = None):
        """``samples()`` before the conversion to Python ints: ``(keys [K, words] uint64 ascending, counts [K] int64)``,
        the form ``dist.merge_histograms`` exchanges between GPUs."""
        from .dist import merge_tables
        max_lit = max((abs(l) for c in self.clauses for l in c), default=0)
        if self.n_vars > max_lit:
            # VariableAssignment(clauses=...) sizes its vector by the largest literal (utils/VariableAssignment.py:34-36);
            # the reference then fails in assign_all_from_bit_list on the first sample -- raised here before any GPU work
            raise IndexError("list assignment index out of range")
        tables = []
        total = sat_total = launched = 0
        still_needed = int(n_samples)
        words = -(-self.n_vars // 64)
        stop = False
        while still_needed > 0 and not stop and (max_chains is None or launched < max_chains):
            chains = self._launch_chains(still_needed, None if max_chains is None else max_chains - launched,
                                         sat_total / total if total else None)
            if self.ctx.graph is not self.unit or self.ctx.chains != chains:
                self.ctx.set_graph(self.unit, chains=chains, group_graphs=self.batch_chains)
                self.model._graph_key = None
            self.ctx.sample_enqueue(diffusion_steps, test_rounds, seed=self.seed,
                                    chain_offset=self.chain_offset + self._chains_consumed + launched)
            is_sat = self.ctx.sample_fetch_sat()
            used, sat_used, stop = consume_batches(is_sat, self.batch_chains, still_needed, total, sat_total,
                                                   self.min_sat_rate)
            if stop:
                print("too many unsat samples; stopping diffusion")
            if sat_used:
                keys, counts, n_sat = self.ctx.hist_reduce(used)        # sort / unique / count on the device
                assert n_sat == sat_used
                tables.append((keys, counts))
            total += used
            sat_total += sat_used
            still_needed -= sat_used
            launched += chains
        self._chains_consumed += launched
        keys, counts = merge_tables(tables, words)
        self.last_stats = {"total": total, "sat": sat_total, "chains_launched": launched, "distinct": int(len(counts))}
        print("success rate: ", sat_total / total if total else 0.0)
        return keys, counts
