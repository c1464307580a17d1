"""``QuerySAT`` with the reference's call signature, running on libdsat (sm_100a CUDA).

Mirror of reference ``model/query_sat.py``: constructor ``:86-89``, ``call`` ``:133-184``,
``diffusion_step`` ``:467-481``, ``predict_step`` ``:424-451`` and the helpers
``randomized_rounding_tf`` ``:55-60``, ``distribution_at_time`` ``:66-68``, ``add_t_emb`` ``:70-74``,
``construct_training_input`` ``:76-82``.  Only the inference branch exists here (``training=True``
raises): training is outside the hot path this package replaces.

Sparse inputs are accepted in any of these forms: an object with ``indices`` and ``dense_shape``
(``tf.SparseTensor``-like), a torch sparse COO tensor, a ``(indices, dense_shape)`` tuple, or — for the
graph membership matrices — a 1-D array of graph ids.
"""

from __future__ import annotations

import hashlib

import numpy as np

from . import _lib
from .graph import unit_graph_from_reference_coo
from .weights import QuerySATWeights, init_weights

t_power = 1 / 2  # reference model/query_sat.py:13


# ------------------------------------------------------------------------ free helper functions
def distribution_at_time(x, time_increment):
    n_classes = 2
    return x * (1 - time_increment) + time_increment / n_classes


def randomized_rounding_tf(x, noise=None, rng=None):
    """``floor(x[...,0:1] + U)`` -> ``[r, 1-r]``; ``noise`` may be injected (the reference draws it)."""
    x = np.asarray(x, dtype=np.float32)
    x0 = x[..., 0:1]
    if noise is None:
        noise = (rng or np.random.default_rng()).random(x0.shape, dtype=np.float32)
    rounded = np.floor(x0 + np.asarray(noise, dtype=np.float32).reshape(x0.shape))
    return np.concatenate([rounded, 1 - rounded], axis=-1).astype(np.float32)


def add_t_emb(both_nums_noisy, noise_scale):
    both = np.asarray(both_nums_noisy, dtype=np.float32)
    t_emb = np.zeros((both.shape[0], 1), dtype=np.float32) + np.float32(noise_scale)
    return np.concatenate([both, t_emb], axis=-1)


def construct_training_input(both_nums_int, noise_scale, rng=None):
    one_hot = np.eye(2, dtype=np.float32)[np.asarray(both_nums_int, dtype=np.int64)]
    num_at_t = distribution_at_time(one_hot, np.power(np.float32(noise_scale), np.float32(t_power)))
    return randomized_rounding_tf(num_at_t, rng=rng)


# ------------------------------------------------------------------------------- input adapters
def _coo(sparse):
    """-> (indices int64 [E,2], dense_shape (rows, cols))"""
    if isinstance(sparse, (tuple, list)) and len(sparse) == 2:
        idx, shape = sparse
        return np.asarray(idx, dtype=np.int64).reshape(-1, 2), (int(shape[0]), int(shape[1]))
    if hasattr(sparse, "indices") and hasattr(sparse, "dense_shape"):
        idx = sparse.indices
        idx = idx.numpy() if hasattr(idx, "numpy") else idx
        shape = sparse.dense_shape
        shape = shape.numpy() if hasattr(shape, "numpy") else shape
        return np.asarray(idx, dtype=np.int64).reshape(-1, 2), (int(shape[0]), int(shape[1]))
    if hasattr(sparse, "is_sparse") and sparse.is_sparse:  # torch sparse COO
        sp = sparse.coalesce()
        return sp.indices().t().cpu().numpy().astype(np.int64), tuple(int(s) for s in sp.shape)
    raise TypeError("expected a sparse matrix as (indices, dense_shape), SparseTensor-like or torch sparse COO")


def _graph_ids(membership, count):
    """Graph id per node from a G x count membership matrix (or a 1-D id array / None)."""
    if membership is None:
        return np.zeros(count, dtype=np.int64), 1
    arr = membership
    if hasattr(arr, "numpy") and not hasattr(arr, "indices"):
        arr = arr.numpy()
    if isinstance(arr, np.ndarray) and arr.ndim == 1:
        return arr.astype(np.int64), int(arr.max()) + 1 if arr.size else 1
    if isinstance(arr, np.ndarray) and arr.ndim == 2:
        return np.argmax(arr, axis=0).astype(np.int64), arr.shape[0]
    idx, shape = _coo(arr)
    ids = np.zeros(count, dtype=np.int64)
    ids[idx[:, 1]] = idx[:, 0]
    return ids, shape[0]


def is_graph_sat(predictions, adj_indices, adj_shape, clause_graph_ids, n_graphs):
    """Per-graph 0/1 flags: every clause of the graph has a true literal under ``round(sigmoid(logits))``
    (reference ``utils/sat.py:165-180``; ``tf.round`` is half-to-even, so the bit is ``sigmoid > 0.5`` strictly)."""
    n, m = adj_shape[0] // 2, adj_shape[1]
    z = np.asarray(predictions, dtype=np.float32).reshape(n)
    bits = ((np.float32(1.0) / (np.float32(1.0) + np.exp(-z, dtype=np.float32))) > np.float32(0.5)).astype(np.float32)
    literals = np.concatenate([bits, 1.0 - bits])
    clause_sat = np.zeros(m, dtype=np.float32)
    np.add.at(clause_sat, adj_indices[:, 1], literals[adj_indices[:, 0]])
    clause_sat = np.clip(clause_sat, 0.0, 1.0)
    sat_in_g = np.zeros(n_graphs, dtype=np.float32)
    np.add.at(sat_in_g, clause_graph_ids, clause_sat)
    total_in_g = np.bincount(clause_graph_ids, minlength=n_graphs).astype(np.float32)
    return np.clip(sat_in_g + 1.0 - total_in_g, 0.0, 1.0)


class QuerySAT:
    def __init__(self, optimizer=None, feature_maps=128, msg_layers=3, vote_layers=3, train_rounds=32,
                 test_rounds=64, query_maps=128, supervised=True, trial=None, *, weights: QuerySATWeights = None,
                 device: int = 0, precision: str = "fp32", seed: int = 0, context=None, **kwargs):
        if trial is not None:
            raise NotImplementedError("optuna trials configure training, which is outside this package")
        if msg_layers != 3 or not supervised:
            raise NotImplementedError("the CUDA path implements the reference defaults msg_layers=3, supervised=True")
        self.optimizer = optimizer          # ignored at inference, kept for signature parity
        self.supervised = supervised
        self.train_rounds = train_rounds
        self.test_rounds = test_rounds
        self.feature_maps = feature_maps
        self.query_maps = query_maps
        self.vote_layers = vote_layers
        self.logit_maps = 8
        self.prediction_tries = 1
        self._rng = np.random.default_rng(seed)
        self._seed = int(seed)
        self._calls = 0
        self.weights = weights if weights is not None else init_weights(feature_maps, query_maps, seed=1234)
        if (self.weights.feature_maps, self.weights.query_maps) != (feature_maps, query_maps):
            raise ValueError("weights were built for feature_maps=%d query_maps=%d" %
                             (self.weights.feature_maps, self.weights.query_maps))
        self.ctx = context if context is not None else _lib.Context(device)
        self.ctx.set_model(self.weights)
        # "fp32" = fp32-accurate Dense layers on the tensor cores (DSAT_F32_TC); widths those kernels do not tile
        # (feature_maps = 256) run the same fp32 arithmetic on the CUDA cores
        try:
            self.ctx.set_precision(_lib.PRECISIONS[precision])
        except _lib.DsatError:
            if _lib.PRECISIONS[precision] != _lib.F32_TC:
                raise
            self.ctx.set_precision(_lib.F32)
        self.precision = precision
        self._graph_key = None

    # ------------------------------------------------------------------------------ weights
    def set_weights(self, weights: QuerySATWeights):
        self.weights = weights
        self.ctx.set_model(weights)

    # -------------------------------------------------------------------------------- graph
    def _bind_graph(self, adj_matrix, clauses_graph, variables_graph):
        idx, shape = _coo(adj_matrix)
        n, m = shape[0] // 2, shape[1]
        vg, g1 = _graph_ids(variables_graph, n)
        cg, g2 = _graph_ids(clauses_graph, m)
        key = hashlib.sha1(idx.tobytes() + vg.tobytes() + cg.tobytes() + repr(shape).encode()).hexdigest()
        if key != self._graph_key:
            unit = unit_graph_from_reference_coo(idx, shape, vg, cg)
            # the whole reference batch is one early-exit group (model/query_sat.py:330-338)
            self.ctx.set_graph(unit, chains=1, group_graphs=0)
            self._graph_key = key
        return n, m

    # --------------------------------------------------------------------------------- call
    def call(self, adj_matrix, clauses_graph=None, variables_graph=None, training=None, labels=None, mask=None,
             noise_scale=None, noisy_num=None, denoised_num=None, *, normals=None):
        """-> ``(last_logits [N,1], loss, step)`` as reference ``:184``.  ``normals`` ([rounds,N,4]) may be
        injected; otherwise they come from the device Philox stream."""
        if training:
            raise NotImplementedError("training is outside the sampling hot path")
        if denoised_num is not None:
            raise NotImplementedError("self-supervised denoised_num input is not used by the sampler (self_supervised=False)")
        n, _ = self._bind_graph(adj_matrix, clauses_graph, variables_graph)
        if noise_scale is None:
            noise_scale = float(self._rng.random())                       # :144
        noise_scale = float(np.float32(noise_scale))
        if labels is None:
            labels = self._rng.integers(0, 2, size=n, dtype=np.int32)     # :145
        labels = np.asarray(labels, dtype=np.int32).reshape(n)
        if noisy_num is None:
            noisy_num = construct_training_input(labels, noise_scale, rng=self._rng)   # :214
        rounds = self.train_rounds if training else self.test_rounds      # :151
        self._calls += 1
        pred, steps, loss = self.ctx.model_call(noise_scale, noisy_num, labels=labels, normals=normals, rounds=rounds,
                                                seed=self._seed + self._calls)
        return pred.reshape(n, 1), np.float32(loss[0]), int(steps[0])

    __call__ = call

    def diffusion_step(self, adj_matrix, clauses_graph, variables_graph, solutions, noise_scale, noisy_num, *,
                       labels=None, normals=None):
        predictions, loss, step = self.call(adj_matrix, clauses_graph, variables_graph, training=False, labels=labels,
                                            noise_scale=noise_scale, noisy_num=noisy_num, normals=normals)
        return {"steps_taken": step, "loss": loss, "prediction": predictions[:, 0]}

    def predict_step(self, adj_matrix, clauses_graph, variables_graph, solutions=None):
        """Reference ``:424-451``: one call, or ``prediction_tries`` calls where every graph keeps the logits of the
        first try that satisfied it (graphs never solved end with zeros, as in the reference)."""
        if self.prediction_tries == 1:
            predictions, loss, step = self.call(adj_matrix, clauses_graph, variables_graph, training=False)
        else:
            idx, shape = _coo(adj_matrix)
            n, m = shape[0] // 2, shape[1]
            vg, n_graphs = _graph_ids(variables_graph, n)
            cg, _ = _graph_ids(clauses_graph, m)
            final = np.zeros((n, 1), dtype=np.float32)
            solved = np.zeros(n_graphs, dtype=np.float32)
            for _ in range(self.prediction_tries):
                predictions, loss, step = self.call(adj_matrix, clauses_graph, variables_graph, training=False)
                sat = np.clip(is_graph_sat(predictions, idx, shape, cg, n_graphs) - solved, 0.0, 1.0)   # newly solved
                final += predictions * sat[vg][:, None]
                solved = solved + sat
            predictions = final
        return {"steps_taken": step, "loss": loss, "prediction": predictions[:, 0]}

    def plot_step(self, adj_matrix, clauses_graph, variables_graph, solutions, noise_scale):
        """Reference ``:453-465``: a call at a given noise level with the solutions as labels."""
        labels = solutions.flat_values if hasattr(solutions, "flat_values") else solutions
        labels = labels.numpy() if hasattr(labels, "numpy") else labels
        labels = np.concatenate([np.asarray(x).reshape(-1) for x in labels]) if isinstance(labels, (list, tuple)) else labels
        predictions, loss, step = self.call(adj_matrix, clauses_graph, variables_graph, training=False, labels=labels,
                                            noise_scale=noise_scale)
        return {"steps_taken": step, "loss": loss, "prediction": predictions[:, 0]}

    def get_config(self):
        return {"model": self.__class__.__name__, "feature_maps": self.feature_maps, "query_maps": self.query_maps,
                "train_rounds": self.train_rounds, "test_rounds": self.test_rounds, "mlp_layers": self.vote_layers}
