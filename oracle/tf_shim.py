"""A torch-backed stand-in for the TensorFlow / Keras / TFP calls the reference's sampling path makes.
TEST INFRASTRUCTURE ONLY (used by tests/golden/make_model_golden.py in the build container).

Purpose: TensorFlow 2.4 cannot be installed offline, but the reference's model is plain Python on top
of ~60 TF calls.  Registering this module as ``tensorflow`` lets the UNMODIFIED reference sources
``model/query_sat.py``, ``model/mlp.py``, ``layers/normalization.py``, ``loss/sat.py``, ``utils/sat.py``,
``metrics/sat_metrics.py`` and ``satuniformity/DiffusionSampler.py`` run here, so their control flow,
tensor plumbing and op order produce golden vectors for the oracle and the CUDA path.  What remains
unpinned is the numerical behaviour of each TF kernel itself: every function below implements the
documented semantics of the TF op of the same name in fp32 torch.

Random ops do not draw: they pop tensors from ``NOISE`` so that the caller controls all randomness.
"""

from __future__ import annotations

import sys
import types

import numpy as np
import torch

float32, int32, int64, bool_ = torch.float32, torch.int32, torch.int64, torch.bool


# ------------------------------------------------------------------------------------------- noise
class NoiseFeed:
    def __init__(self):
        self.normals, self.uniforms, self.labels = [], [], []
        self.log = []

    def clear(self):
        self.normals, self.uniforms, self.labels, self.log = [], [], [], []

    def pop(self, kind, shape):
        queue = getattr(self, kind)
        if not queue:
            raise RuntimeError("tf_shim: the reference asked for %s noise of shape %r but none was queued" % (kind, shape))
        value = queue.pop(0)
        self.log.append((kind, tuple(value.shape)))
        if tuple(int(s) for s in shape) != tuple(value.shape):
            raise RuntimeError("tf_shim: %s noise shape %r requested, %r queued" % (kind, tuple(shape), tuple(value.shape)))
        return value


NOISE = NoiseFeed()


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype if dtype is not None else (float32 if isinstance(x, float) else None))


def _dims(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1)]
    if isinstance(shape, (int, np.integer)):           # tf.ones(n): a scalar shape is a vector length
        return [int(shape)]
    return [int(v) for v in shape]


# ------------------------------------------------------------------------------------------ sparse
class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices = _t(indices, int64).reshape(-1, 2)
        self.values = _t(values)
        self.dense_shape = torch.as_tensor([int(v) for v in dense_shape], dtype=int64)

    @property
    def shape(self):
        return tuple(int(v) for v in self.dense_shape)

    def __truediv__(self, other):
        # sparse / dense[rows,1] broadcast, as tf.sparse division by a column of row sums
        other = _t(other, self.values.dtype)
        if other.dim() == 2 and other.shape[1] == 1:
            return SparseTensor(self.indices, self.values / other[self.indices[:, 0], 0], self.dense_shape)
        return SparseTensor(self.indices, self.values / other, self.dense_shape)


class _Sparse(types.SimpleNamespace):
    SparseTensor = SparseTensor

    @staticmethod
    def transpose(sp):
        idx = sp.indices[:, [1, 0]]
        order = np.lexsort((idx[:, 1].numpy(), idx[:, 0].numpy()))      # canonical row-major order
        order = torch.from_numpy(order)
        return SparseTensor(idx[order], sp.values[order], [sp.shape[1], sp.shape[0]])

    @staticmethod
    def reduce_sum(sp, axis=None, keepdims=False):
        rows, cols = sp.shape
        if axis in (1, -1):
            out = torch.zeros(rows, dtype=sp.values.dtype).index_add_(0, sp.indices[:, 0], sp.values)
            return out[:, None] if keepdims else out
        if axis == 0:
            out = torch.zeros(cols, dtype=sp.values.dtype).index_add_(0, sp.indices[:, 1], sp.values)
            return out[None, :] if keepdims else out
        return sp.values.sum()

    @staticmethod
    def sparse_dense_matmul(sp, dense, adjoint_a=False):
        dense = _t(dense)
        r, c = (sp.indices[:, 1], sp.indices[:, 0]) if adjoint_a else (sp.indices[:, 0], sp.indices[:, 1])
        n_rows = sp.shape[1] if adjoint_a else sp.shape[0]
        vals = sp.values.to(dense.dtype)
        out = torch.zeros(n_rows, dense.shape[1], dtype=dense.dtype)
        return out.index_add_(0, r, dense[c] * vals[:, None])

    @staticmethod
    def to_dense(sp):
        out = torch.zeros(sp.shape, dtype=sp.values.dtype)
        out.index_put_((sp.indices[:, 0], sp.indices[:, 1]), sp.values, accumulate=True)
        return out


# ------------------------------------------------------------------------------------------ autodiff
class GradientTape:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        if isinstance(sources, (list, tuple)):
            return list(torch.autograd.grad(target, list(sources), retain_graph=True, allow_unused=True))
        (g,) = torch.autograd.grad(target, [sources], retain_graph=True)
        return g


class TensorArray:
    def __init__(self, dtype, size=0, dynamic_size=True, clear_after_read=True):
        self.items = {}

    def write(self, index, value):
        self.items[int(index)] = _t(value)
        return self

    def stack(self):
        return torch.stack([self.items[k] for k in sorted(self.items)])


# -------------------------------------------------------------------------------------- keras pieces
class Layer:
    def __init__(self, *a, name=None, **k):
        self.name = name

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class Model(Layer):
    pass


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, **k):
        super().__init__()
        self.units, self.activation, self.use_bias = units, activation, use_bias
        self.kernel = None
        self.bias = None

    def set_weights(self, kernel, bias):
        self.kernel = torch.as_tensor(np.asarray(kernel), dtype=float32).clone().requires_grad_(True)
        self.bias = torch.as_tensor(np.asarray(bias), dtype=float32).clone().requires_grad_(True)
        assert self.kernel.shape[1] == self.units

    def call(self, inputs, training=None):
        if self.kernel is None:
            raise RuntimeError("tf_shim.Dense used before set_weights")
        out = _t(inputs, float32) @ self.kernel
        if self.use_bias:
            out = out + self.bias
        return self.activation(out) if self.activation is not None else out


class Lambda(Layer):
    def __init__(self, fn, **k):
        super().__init__()
        self.fn = fn

    def call(self, x, **k):
        return self.fn(x)


class _Mean:
    def update_state(self, *a, **k):
        pass

    def reset_states(self):
        pass

    def result(self):
        return torch.tensor(0.0)


# ----------------------------------------------------------------------------------------- tfp piece
class Bernoulli:
    def __init__(self, probs=None, logits=None):
        self.probs = _t(probs, float32)

    def kl_divergence(self, other):
        pa, pb = torch.broadcast_tensors(self.probs, other.probs)
        qa = 1 - pa
        t1 = torch.where(pa == 0, torch.zeros_like(pa), pa * (torch.log(pa) - torch.log(pb)))
        t2 = torch.where(qa == 0, torch.zeros_like(pa), qa * (torch.log1p(-pa) - torch.log1p(-pb)))
        return t1 + t2


# ----------------------------------------------------------------------------------------- tf.* ops
def function(fn=None, **kwargs):
    if fn is not None and callable(fn):
        return fn
    return lambda f: f


def shape(x):
    if isinstance(x, SparseTensor):
        return [int(v) for v in x.dense_shape]
    return torch.as_tensor(list(_t(x).shape), dtype=int64) if False else _ShapeList(_t(x).shape)


class _ShapeList(list):
    """list of ints that also supports slicing like a tensor (tf.shape(x)[0:1])"""

    def __init__(self, dims):
        super().__init__(int(d) for d in dims)


def ones(shp, dtype=float32):
    return torch.ones(_dims(shp), dtype=dtype)


def zeros(shp, dtype=float32):
    return torch.zeros(_dims(shp), dtype=dtype)


def _random_uniform(shp, minval=0, maxval=None, dtype=float32):
    dims = _dims(shp) if not isinstance(shp, tuple) or len(shp) else []
    if dtype in (int32, int64):
        return NOISE.pop("labels", dims).to(dtype)
    return NOISE.pop("uniforms", dims).to(float32)


def _random_normal(shp, mean=0.0, stddev=1.0, dtype=float32):
    return NOISE.pop("normals", _dims(shp)).to(float32)


def reshape(x, shp):
    return _t(x).reshape(_dims(shp))


def maximum(a, b):
    a = _t(a)
    return torch.maximum(a, _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b.to(a.dtype))


def minimum(a, b):
    a = _t(a)
    return torch.minimum(a, _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b.to(a.dtype))


def concat(values, axis=0):
    return torch.cat([_t(v) for v in values], dim=axis)


def split(value, num, axis=0):
    return torch.chunk(_t(value), num, dim=axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def tile(x, multiples):
    return _t(x).repeat(*_dims(multiples))


def cast(x, dtype):
    if isinstance(x, SparseTensor):
        return SparseTensor(x.indices, x.values.to(dtype), x.dense_shape)
    return _t(x).to(dtype)


def sort(x, axis=-1, direction="ASCENDING"):
    return torch.sort(_t(x), dim=axis, descending=(direction == "DESCENDING")).values


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_min(x, axis=None):
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis).values


def argmin(x, axis=None, output_type=int64):
    return torch.argmin(_t(x), dim=axis).to(output_type)


def gather(params, indices, batch_dims=0, axis=None):
    params, indices = _t(params), _t(indices, int64)
    if batch_dims == 1:
        return torch.gather(params, 1, indices[:, None])[:, 0]
    return params[indices]


def stop_gradient(x):
    return _t(x).detach()


def one_hot(indices, depth, dtype=float32):
    return torch.nn.functional.one_hot(_t(indices, int64), depth).to(dtype)


def clip_by_value(x, lo, hi):
    return torch.clamp(_t(x), lo, hi)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def tf_range(*args, dtype=None):
    if dtype is not None:
        return torch.arange(*args, dtype=dtype)
    return range(*[int(a) for a in args])


def _pow(x, y):
    return torch.pow(_t(x, float32), y)


def install():
    """Register the shim as `tensorflow`, `tensorflow_probability`, … in sys.modules. Returns the tf module."""
    # EagerTensor.numpy() works on anything; torch refuses on tensors that are part of an autograd graph
    # (the Dense kernels require grad so that the reference's inner GradientTape works)
    if not getattr(torch.Tensor.numpy, "_dsat_patched", False):
        _orig_numpy = torch.Tensor.numpy

        def _numpy(self, *a, **k):
            return _orig_numpy(self.detach(), *a, **k)

        _numpy._dsat_patched = True
        torch.Tensor.numpy = _numpy

    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.int64, tf.bool = float32, int32, int64, bool_
    tf.Tensor = torch.Tensor
    tf.SparseTensor = SparseTensor
    tf.RaggedTensor = object
    tf.sparse = _Sparse()
    tf.GradientTape = GradientTape
    tf.TensorArray = TensorArray
    tf.function = function
    for spec in ("SparseTensorSpec", "RaggedTensorSpec", "TensorSpec"):
        setattr(tf, spec, lambda *a, **k: None)
    tf.shape, tf.ones, tf.zeros, tf.reshape = shape, ones, zeros, reshape
    tf.maximum, tf.minimum, tf.concat, tf.split = maximum, minimum, concat, split
    tf.expand_dims, tf.squeeze, tf.tile, tf.cast, tf.sort = expand_dims, squeeze, tile, cast, sort
    tf.reduce_sum, tf.reduce_mean, tf.reduce_min, tf.argmin = reduce_sum, reduce_mean, reduce_min, argmin
    tf.gather, tf.stop_gradient, tf.one_hot, tf.clip_by_value, tf.stack = gather, stop_gradient, one_hot, clip_by_value, stack
    tf.range = tf_range
    tf.repeat = lambda x, repeats, axis=None: torch.repeat_interleave(_t(x), _t(repeats, int64), dim=axis)
    tf.round = lambda x: torch.round(_t(x))                      # half-to-even, like tf.round
    tf.floor = lambda x: torch.floor(_t(x))
    tf.sigmoid = lambda x: torch.sigmoid(_t(x))
    tf.exp = lambda x: torch.exp(_t(x))
    tf.square = lambda x: torch.square(_t(x))
    tf.sqrt = lambda x: torch.sqrt(_t(x, float32))
    tf.abs = lambda x: torch.abs(_t(x))
    tf.sign = lambda x: torch.sign(_t(x))
    tf.equal = lambda a, b: _t(a) == _t(b)
    tf.transpose = lambda x: _t(x).t()
    tf.convert_to_tensor = lambda x, dtype=None: _t(x, dtype)
    tf.constant_initializer = lambda v: v
    tf.math = types.SimpleNamespace(rsqrt=lambda x: torch.rsqrt(_t(x)), pow=_pow, log=lambda x: torch.log(_t(x)),
                                    segment_sum=None)
    tf.nn = types.SimpleNamespace(softplus=lambda x: torch.nn.functional.softplus(_t(x)),
                                  leaky_relu=lambda x, alpha=0.2: torch.nn.functional.leaky_relu(_t(x), alpha),
                                  sigmoid=lambda x: torch.sigmoid(_t(x)))
    tf.random = types.SimpleNamespace(uniform=_random_uniform, normal=_random_normal)
    tf.summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)
    tf.metrics = types.SimpleNamespace(Mean=_Mean)
    tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=Layer, Dropout=None),
                                     models=types.SimpleNamespace(Model=Model))
    tf.data = types.SimpleNamespace(experimental=types.SimpleNamespace(AUTOTUNE=-1))

    def reg(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    sys.modules["tensorflow"] = tf
    reg("tensorflow.keras")
    reg("tensorflow.keras.models", Model=Model)
    reg("tensorflow.keras.optimizers", Optimizer=object)
    reg("tensorflow.python")
    reg("tensorflow.python.keras")
    reg("tensorflow.python.keras.layers", Dense=Dense, Lambda=Lambda)
    reg("tensorflow_probability", distributions=types.SimpleNamespace(Bernoulli=Bernoulli))
    reg("optuna", Trial=object)
    hp_names = ["HP_MODEL", "HP_FEATURE_MAPS", "HP_QUERY_MAPS", "HP_TRAIN_ROUNDS", "HP_TEST_ROUNDS", "HP_MLP_LAYERS",
                "HP_TRAINABLE_PARAMS", "HP_TASK"]
    reg("utils.parameters_log", **{n: n for n in hp_names}, __all__=hp_names)
    reg("pysat")
    reg("pysat.formula", CNF=object)
    reg("pysat.solvers", Glucose4=object)
    return tf
