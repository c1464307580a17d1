"""Philox4x32-10 noise streams, numpy restatement of ``csrc/dsat_common.cuh`` (host side).

The kernels draw all randomness of a run from a counter-based generator keyed by ``seed``; the counter is
``(element lo, element hi, step<<16 | round, stream)`` with ``element = global_chain * n_unit_vars + var``.
Results therefore do not depend on how chains are split over launches or GPUs.  The reference draws
the same quantities with ``tf.random`` (``model/query_sat.py:57,145,239``); TensorFlow's stateful
streams cannot be reproduced, so noise-matched parity runs inject tensors generated here.
"""

from __future__ import annotations

import numpy as np

STREAM_NORMAL, STREAM_UNIFORM, STREAM_LABEL = 0, 1, 2
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _unit_float(x):
    bits = (x & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)
    return bits.view(np.float32) - np.float32(1.0)


def _draw(seed, elements, step, rnd, stream):
    elements = np.asarray(elements, dtype=np.uint64)
    c0 = (elements & _MASK).astype(np.uint32)
    c1 = (elements >> np.uint64(32)).astype(np.uint32)
    c2 = np.uint32(((int(step) << 16) | int(rnd)) & 0xFFFFFFFF)
    return philox4x32_10(c0, c1, c2, np.uint32(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def uniforms(seed, elements, step):
    return _unit_float(_draw(seed, elements, step, 0, STREAM_UNIFORM)[0])


def labels(seed, elements, step):
    return (_draw(seed, elements, step, 0, STREAM_LABEL)[0] & np.uint32(1)).astype(np.int32)


def normals(seed, elements, step, rnd):
    """[len(elements), 4] float32, Box-Muller on (x,y) and (z,w) like the device code."""
    x, y, z, w = _draw(seed, elements, step, rnd, STREAM_NORMAL)

    def bm(a, b):
        u1 = np.maximum(_unit_float(a), np.float32(1.0e-7))
        ang = np.float32(6.283185307179586) * _unit_float(b)
        rad = np.sqrt(np.float32(-2.0) * np.log(u1))
        return np.sin(ang) * rad, np.cos(ang) * rad

    n0, n1 = bm(x, y)
    n2, n3 = bm(z, w)
    return np.stack([n0, n1, n2, n3], axis=-1).astype(np.float32)
