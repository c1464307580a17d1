"""Weights of the five QuerySAT MLPs (twelve Dense layers) and their container.

Layer shapes follow reference ``model/query_sat.py:101-122`` and ``model/mlp.py:13-40``
(``do_layer_norm=False`` branch: ``layer_count-1`` hidden Dense with leaky_relu, one linear
output Dense; Keras defaults glorot-uniform kernel, zero bias; ``y = x @ W + b`` with
``W[in, out]``).

Input column order of each first layer (what the kernels' weight re-packing relies on):

* ``variables_query`` / ``lit_query``: ``v1 = [variables(F) | normal(4) | noisy(2) | noise_scale(1) | denoised(2)]``
  (``model/query_sat.py:214-219,239``)
* ``clause_update``: ``[clause_state(F) | clause_messages(Q) | 4*clauses_loss(Q)]`` (``:258``)
* ``update_gate``: ``[variables_grad(Q) | v1(F+9) | loss_pos(Q) | loss_neg(Q)]`` (``:277``)
* ``variables_output``: ``variables(F)`` (``:283``)

``model_path`` is either a ``.npz`` written by :func:`save_weights` or a TensorFlow checkpoint directory /
prefix as the reference writes it (read without TensorFlow by :mod:`diffusionsat_b200.tf_checkpoint`,
SURVEY.md section 8f item 1).
"""

from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

AUX_WIDTH = 9          # normal(4) + noisy(2) + noise_scale(1) + denoised(2)
LOGIT_MAPS = 8         # reference model/query_sat.py:99
MLP_ORDER = ("variables_query", "lit_query", "clause_update", "update_gate", "variables_output")


def mlp_layer_dims(feature_maps: int = 128, query_maps: int = 128):
    """``{mlp_name: [(in, out), ...]}`` in the order the layers are applied."""
    f, q = int(feature_maps), int(query_maps)
    v1 = f + AUX_WIDTH

    def chain(n_in, hidden, n_out, layers):
        dims, cur = [], n_in
        for _ in range(layers - 1):
            dims.append((cur, hidden))
            cur = hidden
        dims.append((cur, n_out))
        return dims

    return OrderedDict([
        ("variables_query", chain(v1, int(q * 1.2), q, 2)),            # :119  query_layers=2, query_scale=1.2
        ("lit_query", chain(v1, q * 4, q * 2, 3)),                     # :122  msg_layers=3
        ("clause_update", chain(f + 2 * q, int(f * 1.6), f + q, 2)),   # :120  clauses_layers=2, clauses_scale=1.6
        ("update_gate", chain(q + v1 + 2 * q, int(f * 1.8), f, 3)),    # :117  update_layers=3, update_scale=1.8
        ("variables_output", chain(f, int(f * 1), LOGIT_MAPS, 2)),     # :118  output_layers=2, output_scale=1
    ])


def flat_layer_names(feature_maps: int = 128, query_maps: int = 128):
    names = []
    for mlp, dims in mlp_layer_dims(feature_maps, query_maps).items():
        for i in range(len(dims)):
            names.append("%s/%d" % (mlp, i))
    return names


class QuerySATWeights:
    """Ordered ``{"mlp/i": (kernel[in,out] f32, bias[out] f32)}``."""

    def __init__(self, layers: "OrderedDict[str, tuple]", feature_maps=128, query_maps=128):
        self.layers = layers
        self.feature_maps = int(feature_maps)
        self.query_maps = int(query_maps)

    def mlp(self, name):
        n = len(mlp_layer_dims(self.feature_maps, self.query_maps)[name])
        return [self.layers["%s/%d" % (name, i)] for i in range(n)]

    def n_params(self) -> int:
        return int(sum(k.size + b.size for k, b in self.layers.values()))

    def as_dtype(self, dtype):
        out = OrderedDict((k, (w.astype(dtype), b.astype(dtype))) for k, (w, b) in self.layers.items())
        return QuerySATWeights(out, self.feature_maps, self.query_maps)


def init_weights(feature_maps=128, query_maps=128, seed=1234, bias_scale=0.0) -> QuerySATWeights:
    """Seeded glorot-uniform kernels; zero biases unless ``bias_scale`` > 0 (tests use non-zero
    biases so that a missing bias add cannot pass unnoticed)."""
    rng = np.random.default_rng(seed)
    layers = OrderedDict()
    for mlp, dims in mlp_layer_dims(feature_maps, query_maps).items():
        for i, (n_in, n_out) in enumerate(dims):
            limit = np.sqrt(6.0 / (n_in + n_out))
            kernel = rng.uniform(-limit, limit, size=(n_in, n_out)).astype(np.float32)
            bias = (rng.standard_normal(n_out) * bias_scale).astype(np.float32)
            layers["%s/%d" % (mlp, i)] = (kernel, bias)
    return QuerySATWeights(layers, feature_maps, query_maps)


def save_weights(path: str, weights: QuerySATWeights) -> None:
    blobs = {"feature_maps": np.int64(weights.feature_maps), "query_maps": np.int64(weights.query_maps)}
    for name, (kernel, bias) in weights.layers.items():
        blobs[name + "/kernel"] = kernel
        blobs[name + "/bias"] = bias
    np.savez(path, **blobs)


def load_weights(path: str) -> QuerySATWeights:
    """Load a ``.npz`` (a directory is searched for the newest ``*.npz``) or a TensorFlow checkpoint
    (directory with a ``checkpoint`` state file / ``*.index``, or a ``ckpt-N`` prefix). Raises
    ``FileNotFoundError`` when nothing is there; the caller decides about random init, as the
    reference does (``satuniformity/DiffusionSampler.py:221-225``)."""
    from . import tf_checkpoint
    if not (os.path.isdir(path) and any(f.endswith(".npz") for f in os.listdir(path))) and not path.endswith(".npz") \
            and tf_checkpoint.is_tf_checkpoint(path):
        return tf_checkpoint.load_querysat_weights(path)
    if os.path.isdir(path):
        cands = sorted((os.path.getmtime(os.path.join(path, f)), os.path.join(path, f))
                       for f in os.listdir(path) if f.endswith(".npz"))
        if not cands:
            raise FileNotFoundError("no .npz weights under %s" % path)
        path = cands[-1][1]
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    with np.load(path) as blobs:
        f, q = int(blobs["feature_maps"]), int(blobs["query_maps"])
        layers = OrderedDict()
        for name in flat_layer_names(f, q):
            layers[name] = (blobs[name + "/kernel"].astype(np.float32), blobs[name + "/bias"].astype(np.float32))
    for name, (n_in, n_out) in zip(flat_layer_names(f, q),
                                   [d for dims in mlp_layer_dims(f, q).values() for d in dims]):
        if layers[name][0].shape != (n_in, n_out) or layers[name][1].shape != (n_out,):
            raise ValueError("layer %s has shape %r, expected (%d,%d)" % (name, layers[name][0].shape, n_in, n_out))
    return QuerySATWeights(layers, f, q)
