"""CPU tests of the host-side graph build in libdsat (``dsat_graph_build``, include/dsat.h) against the numpy construction
of ``diffusionsat_b200/graph.py`` -- which is itself pinned against the reference's ``compute_adj_indices``
(tests/test_host_golden.py).  Index work: every array must be bit-identical."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from diffusionsat_b200 import dist as D
from diffusionsat_b200 import graph as G
from diffusionsat_b200 import synth

ARRAYS = ("cl_rowptr", "cl_lit", "lit_rowptr", "lit_clause", "var_seg", "clause_seg")


def assert_same_graph(a, b):
    assert (a.n_vars, a.n_clauses, a.nnz, a.n_graphs) == (b.n_vars, b.n_clauses, b.nnz, b.n_graphs)
    for name in ARRAYS:
        x, y = getattr(a, name), getattr(b, name)
        assert x.dtype == y.dtype == np.int32, name
        np.testing.assert_array_equal(x, y, err_msg=name)


def test_native_build_equals_numpy_on_mixed_ksat(dsat_lib):
    for seed in range(12):
        n = 3 + 9 * seed
        n_vars, clauses = synth.random_ksat_mixed(n, max(1, int(4.3 * n)), seed=seed)
        assert_same_graph(G.build_unit_graph(n_vars, clauses, native=True), G.build_unit_graph(n_vars, clauses, native=False))


def test_native_build_edge_cases(dsat_lib):
    cases = [
        (4, [[1, 1, -1], [], [3, -3, 3, 2], [-2, -2]]),        # repeated literals, an empty clause, an unused variable
        (1, [[1]]),
        (2, [[], []]),                                         # no edge at all
        (3, [[-3, 2, -1, 1, 3, -2]]),                          # one clause holding every literal
        (5, [[5], [-5], [5, -5]]),                             # only the last variable occurs
    ]
    for n_vars, clauses in cases:
        a = G.build_unit_graph(n_vars, clauses, native=True)
        assert_same_graph(a, G.build_unit_graph(n_vars, clauses, native=False))
        assert a.clauses == [list(c) for c in clauses]
    g = G.build_unit_graph(4, cases[0][1], native=True)
    # positives by variable, then negatives (reference literal rows); clause ids ascending with repeats
    assert g.cl_lit.tolist() == [0, 0, 1, 2, 4, 4, 5, 3, 3]
    assert g.lit_clause.tolist() == [0, 0, 0, 2, 3, 3, 2, 2, 2]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 12).flatmap(lambda n: st.tuples(
    st.just(n), st.lists(st.lists(st.integers(1, n).flatmap(lambda v: st.sampled_from([v, -v])), max_size=7), max_size=20))))
def test_native_build_equals_numpy_property(dsat_lib, case):
    n_vars, clauses = case
    assert_same_graph(G.build_unit_graph(n_vars, clauses, native=True), G.build_unit_graph(n_vars, clauses, native=False))


def test_union_of_flat_formulas_equals_union_of_lists(dsat_lib):
    rng = np.random.default_rng(0)
    formulas = [synth.random_ksat_mixed(int(rng.integers(3, 60)), int(rng.integers(1, 200)), seed=100 + i) for i in range(40)]
    flat = [G.flatten_formula(*f) for f in formulas]
    want = G.build_union_graph(formulas, native=False)
    assert_same_graph(G.build_union_graph(flat, native=True), want)
    assert_same_graph(G.build_union_graph(flat, native=False), want)
    assert_same_graph(G.build_union_graph(formulas, native=True), want)
    assert want.n_graphs == 40 and want.var_seg[-1] == sum(f[0] for f in formulas)
    # a FlatFormula stands in for the (n_vars, clauses) pair wherever the batching code looks at one
    assert D.pack_batches(flat, 2500) == D.pack_batches(formulas, 2500)
    n_vars, clauses = flat[3]
    assert n_vars == formulas[3][0] and len(clauses) == len(formulas[3][1]) and flat[3][0] == n_vars
    # the union of nothing
    empty = G.build_union_graph([])
    assert empty.n_vars == 0 and empty.n_clauses == 0 and empty.n_graphs == 0


@pytest.mark.parametrize("native", [True, False])
def test_out_of_range_literal_is_rejected_by_both_paths(dsat_lib, native):
    with pytest.raises(ValueError, match="clause 1"):
        G.build_unit_graph(3, [[1, 2], [4, -1]], native=native)
    with pytest.raises(ValueError, match="clause 0"):
        G.build_unit_graph(3, [[0]], native=native)
    with pytest.raises(ValueError):
        G.flatten_formula(2, [[1, -3]])


def test_dsat_graph_build_validates_its_arguments(dsat_lib):
    ip = C.POINTER(C.c_int32)
    lens = np.array([2, 1], dtype=np.int32)
    flat = np.array([1, -2, 3], dtype=np.int32)
    out = [np.zeros(k, dtype=np.int32) for k in (3, 3, 7, 3)]
    bad = C.c_int32(7)

    def call(n_vars, n_clauses, nnz):
        return dsat_lib.dsat_graph_build(n_vars, n_clauses, nnz, lens.ctypes.data_as(ip), flat.ctypes.data_as(ip),
                                         *[a.ctypes.data_as(ip) for a in out], C.byref(bad))
    assert call(3, 2, 3) == 0 and bad.value == -1
    assert out[0].tolist() == [0, 2, 3] and out[1].tolist() == [0, 3, 4] and out[3].tolist() == [0, 0, 1]
    assert call(3, 2, 2) != 0                  # lens do not add up to nnz
    assert call(2, 2, 3) != 0 and bad.value == 1   # literal 3 with two variables: clause 1
    assert call(-1, 2, 3) != 0
