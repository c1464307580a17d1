"""Clause-literal bipartite graph of a CNF formula as CSR/CSC index arrays.

Host side of the hot path. Restates,
without TensorFlow, what the reference builds in

* ``data/dimac.py:14-18``          ``compute_adj_indices``  (COO pairs, positives then negatives)
* ``data/SatSpecifics.py:21-69``   ``create_adj_matrices``  (negative literal rows offset by the
                                    batch-total variable count; graph membership matrices)
* ``data/dimac.py:213-260``        ``prepare_example``      (formulas of a batch are shifted by the
                                    running variable offset: graph g owns variables [off_g, off_g+n_g))
* ``model/query_sat.py:193-197``   degree weights
* ``data/dimac.py:172-174,267-293`` node count ``2n+m`` and the ``max_nodes_per_batch`` packing rule

Layout handed to the CUDA kernels (``include/dsat.h: dsat_set_graph``): one
*unit* graph (a formula, or a disjoint union of formulas) described once and
shared by all ``n_chains`` replicas. A literal is coded ``2*var + sign`` with
``var`` 0-based inside the unit and ``sign`` 1 for a negated literal; the
reference's literal row of the same literal is ``sign*N_total + chain*n + var``.

* ``cl_rowptr[m+1], cl_lit[nnz]``    clause -> literal codes, duplicates kept, ordered inside a
  clause by the reference literal row (positives by variable, then negatives), i.e. the order in
  which ``tf.sparse.transpose(adj_matrix)`` is stored (``model/query_sat.py:188``)
* ``lit_rowptr[2n+1], lit_clause[nnz]`` literal code -> clause ids, ascending, duplicates kept
  (storage order of ``adj_matrix`` itself for one literal row)
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

MAX_NODES_PER_BATCH = 20000  # reference config.py:35


def compute_adj_indices(clauses):
    """COO (variable, clause) pairs of positive and of negative occurrences.

    Same contract as reference ``data/dimac.py:14-18``: each list is ordered by
    clause index, then by position inside the clause; variables are 0-based;
    repeated literals are repeated pairs.
    """
    pos, neg = [], []
    for idx, clause in enumerate(clauses):
        for lit in clause:
            if lit > 0:
                pos.append([lit - 1, idx])
            elif lit < 0:
                neg.append([-lit - 1, idx])
    return pos, neg


def sat_node_count(n_vars: int, n_clauses: int) -> int:
    """Reference ``data/dimac.py:172-174``."""
    return 2 * n_vars + n_clauses


def chains_per_reference_batch(n_vars: int, n_clauses: int,
                               max_nodes_per_batch: int = MAX_NODES_PER_BATCH) -> int:
    """How many identical copies the reference packs into one batch.

    Greedy rule of ``data/dimac.py:280-287``: copies are appended while the node
    total stays <= max_nodes_per_batch; the first copy always fits.
    """
    nodes = sat_node_count(n_vars, n_clauses)
    return max(1, max_nodes_per_batch // max(nodes, 1))


def _rsqrt_clamped(count: np.ndarray) -> np.ndarray:
    c = np.maximum(count.astype(np.float32), np.float32(1.0))
    return (np.float32(1.0) / np.sqrt(c)).astype(np.float32)


@dataclass
class UnitGraph:
    """Index arrays of one unit (formula or disjoint union of formulas)."""

    n_vars: int
    n_clauses: int
    cl_rowptr: np.ndarray          # int32 [m+1]
    cl_lit: np.ndarray             # int32 [nnz]   literal codes 2*var+sign
    lit_rowptr: np.ndarray         # int32 [2n+1]  indexed by literal code
    lit_clause: np.ndarray         # int32 [nnz]
    var_seg: np.ndarray            # int32 [G+1]   graph g owns variables [var_seg[g], var_seg[g+1])
    clause_seg: np.ndarray         # int32 [G+1]
    _clauses: list = field(default=None, repr=False)                # signed-literal lists, built on first use
    _flat: np.ndarray = field(default=None, repr=False)             # the same literals laid end to end, clause by clause
    _lens: np.ndarray = field(default=None, repr=False)

    @property
    def clauses(self) -> list:
        """Clauses as lists of signed 1-based literals in their original order (only tests and ``reference_coo`` need them)."""
        if self._clauses is None:
            ptr = np.concatenate([[0], np.cumsum(self._lens)])
            flat = self._flat.tolist()
            self._clauses = [flat[ptr[j]:ptr[j + 1]] for j in range(len(self._lens))]
        return self._clauses

    @property
    def nnz(self) -> int:
        return int(self.cl_lit.shape[0])

    @property
    def n_graphs(self) -> int:
        return int(self.var_seg.shape[0] - 1)

    # degree weights, reference model/query_sat.py:193-197 -------------------
    def lit_degree(self) -> np.ndarray:
        return np.diff(self.lit_rowptr).astype(np.int32)

    def degree_weight(self) -> np.ndarray:
        """rsqrt(max(deg(lit),1)) per literal code."""
        return _rsqrt_clamped(self.lit_degree())

    def var_degree_weight(self) -> np.ndarray:
        """4*rsqrt(max(deg(+v)+deg(-v),1)) per variable."""
        deg = self.lit_degree()
        return (np.float32(4.0) * _rsqrt_clamped(deg[0::2] + deg[1::2])).astype(np.float32)

    def rev_degree_weight(self) -> np.ndarray:
        """rsqrt(max(|clause|,1)) per clause."""
        return _rsqrt_clamped(np.diff(self.cl_rowptr))

    # reference-layout COO of `chains` replicas --------------------------------
    def reference_coo(self, chains: int = 1):
        """(literal_row, clause) pairs of ``adj_matrix`` for ``chains`` copies, in the
        reference storage order (all positive pairs, then all negative pairs;
        ``data/SatSpecifics.py:24-35``). Returns int64 [E_total, 2] and the dense shape."""
        n, m = self.n_vars, self.n_clauses
        batched = []
        for c in range(chains):
            off = c * n
            for clause in self.clauses:
                batched.append([lit + off if lit > 0 else lit - off for lit in clause])
        pos, neg = compute_adj_indices(batched)
        n_total = chains * n
        pos = np.asarray(pos, dtype=np.int64).reshape(-1, 2)
        neg = np.asarray(neg, dtype=np.int64).reshape(-1, 2)
        neg[:, 0] += n_total
        return np.concatenate([pos, neg], axis=0), (2 * n_total, chains * m)


def _native_library():
    """libdsat.so if it has been built (its host-side ``dsat_graph_build`` needs no GPU), else None."""
    try:
        from . import _lib
        return _lib.load_library()
    except (RuntimeError, OSError):
        return None


def _graph_arrays_numpy(n: int, lens: np.ndarray, flat: np.ndarray):
    """The four index arrays by two stable argsorts over all edges (the specification ``dsat_graph_build`` is tested against)."""
    m = int(lens.shape[0])
    lens = lens.astype(np.int64)
    flat = flat.astype(np.int64)
    cl_rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=cl_rowptr[1:])
    var = np.abs(flat) - 1
    sign = (flat < 0).astype(np.int64)
    clause_of_edge = np.repeat(np.arange(m, dtype=np.int64), lens)
    # inside a clause: reference literal-row order (positives by variable, then negatives), stable for repeated literals
    order = np.argsort((clause_of_edge * 2 + sign) * max(n, 1) + var, kind="stable")
    cl_lit = (2 * var + sign)[order]
    # literal -> clauses: stable counting sort by literal code keeps clause ids ascending
    order = np.argsort(cl_lit, kind="stable")
    lit_clause = clause_of_edge[order]
    lit_rowptr = np.zeros(2 * n + 1, dtype=np.int64)
    np.cumsum(np.bincount(cl_lit, minlength=2 * n), out=lit_rowptr[1:])
    return (cl_rowptr.astype(np.int32), cl_lit.astype(np.int32), lit_rowptr.astype(np.int32), lit_clause.astype(np.int32))


def _graph_arrays_native(lib, n: int, lens: np.ndarray, flat: np.ndarray):
    """The same arrays from ``dsat_graph_build`` (include/dsat.h): linear in the number of edges."""
    import ctypes as C
    m, nnz = int(lens.shape[0]), int(flat.shape[0])
    lens32 = np.ascontiguousarray(lens, dtype=np.int32)
    flat32 = np.ascontiguousarray(flat, dtype=np.int32)
    cl_rowptr = np.empty(m + 1, dtype=np.int32)
    cl_lit = np.empty(nnz, dtype=np.int32)
    lit_rowptr = np.empty(2 * n + 1, dtype=np.int32)
    lit_clause = np.empty(nnz, dtype=np.int32)
    bad = C.c_int32(-1)
    ip = C.POINTER(C.c_int32)
    rc = lib.dsat_graph_build(n, m, nnz, *[a.ctypes.data_as(ip) for a in (lens32, flat32, cl_rowptr, cl_lit, lit_rowptr,
                                                                           lit_clause)], C.byref(bad))
    if rc != 0:
        raise ValueError("dsat_graph_build failed (%d), clause %d" % (rc, bad.value))
    return cl_rowptr, cl_lit, lit_rowptr, lit_clause


def _graph_from_flat(n: int, lens: np.ndarray, flat: np.ndarray, var_seg, clause_seg, clauses, native=None) -> UnitGraph:
    """CSR/CSC arrays from the literals of all clauses laid end to end (``lens[j]`` literals for clause j).

    ``native``: None = the library's ``dsat_graph_build`` when libdsat.so is built, numpy otherwise; True / False force one."""
    m = int(lens.shape[0])
    if 2 * n + 1 >= 2 ** 31 or int(flat.shape[0]) >= 2 ** 31 or m + 1 >= 2 ** 31:
        raise ValueError("graph too large for 32-bit index arrays (%d variables, %d clauses, %d literals)" % (n, m, flat.shape[0]))
    if flat.size and (np.any(flat == 0) or np.any(np.abs(flat) > n)):
        rowptr = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)])
        bad = int(np.flatnonzero((flat == 0) | (np.abs(flat) > n))[0])
        j = int(np.searchsorted(rowptr, bad, side="right") - 1)
        raise ValueError("literal out of range in clause %d: %r" % (j, flat[rowptr[j]:rowptr[j + 1]].tolist()))
    lib = _native_library() if native is None or native else None
    if native and lib is None:
        raise RuntimeError("libdsat.so is not built: no native graph build")
    if lib is not None:
        cl_rowptr, cl_lit, lit_rowptr, lit_clause = _graph_arrays_native(lib, n, lens, flat)
    else:
        cl_rowptr, cl_lit, lit_rowptr, lit_clause = _graph_arrays_numpy(n, lens, flat)
    if var_seg is None:
        var_seg = [0, n]
    if clause_seg is None:
        clause_seg = [0, m]
    return UnitGraph(
        n_vars=n, n_clauses=m, cl_rowptr=cl_rowptr, cl_lit=cl_lit, lit_rowptr=lit_rowptr, lit_clause=lit_clause,
        var_seg=np.asarray(var_seg, dtype=np.int32), clause_seg=np.asarray(clause_seg, dtype=np.int32),
        _clauses=clauses, _flat=flat, _lens=lens,
    )


def _flatten(clauses):
    m = len(clauses)
    lens = np.fromiter(map(len, clauses), dtype=np.int64, count=m)
    flat = np.fromiter(itertools.chain.from_iterable(clauses), dtype=np.int64, count=int(lens.sum()))
    return lens, flat


@dataclass
class FlatFormula:
    """A formula with its clauses already laid end to end (what a data loader keeps per formula, so that packing it into
    many batches never walks Python lists again; the reference keeps pre-tensorised TFRecords, ``data/dimac.py:129-211``)."""

    n_vars: int
    lens: np.ndarray      # int32 [m]    literals per clause
    flat: np.ndarray      # int32 [nnz]  signed 1-based literals, clause by clause

    @property
    def n_clauses(self) -> int:
        return int(self.lens.shape[0])

    def __iter__(self):                     # unpacks like the (n_vars, clauses) pair it stands for
        yield self.n_vars
        yield self

    def __len__(self):                      # len(clauses)
        return self.n_clauses

    def __getitem__(self, i):               # formula[0] = n_vars, formula[1] = the clauses
        return (self.n_vars, self)[i]


def flatten_formula(n_vars: int, clauses) -> FlatFormula:
    lens, flat = _flatten(clauses)
    if flat.size and (np.any(flat == 0) or np.any(np.abs(flat) > int(n_vars))):
        raise ValueError("literal out of range in a formula with %d variables" % int(n_vars))
    return FlatFormula(int(n_vars), lens.astype(np.int32), flat.astype(np.int32))


def build_unit_graph(n_vars: int, clauses, var_seg=None, clause_seg=None, native=None) -> UnitGraph:
    """CSR/CSC arrays of one formula (or of a union, when the segments are given)."""
    lens, flat = _flatten(clauses)
    return _graph_from_flat(int(n_vars), lens, flat, var_seg, clause_seg, None, native)


def build_union_graph(formulas, native=None) -> UnitGraph:
    """Disjoint union of ``[(n_vars, clauses), ...]`` (or ``FlatFormula`` items) with the reference's variable shift
    (``data/dimac.py:165-170,239-241``): one unit whose graphs are the formulas."""
    flats = [f if isinstance(f, FlatFormula) else flatten_formula(*f) for f in formulas]
    if not flats:
        return _graph_from_flat(0, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), [0], [0], None, native)
    n_per = np.fromiter((f.n_vars for f in flats), dtype=np.int64, count=len(flats))
    m_per = np.fromiter((f.lens.shape[0] for f in flats), dtype=np.int64, count=len(flats))
    e_per = np.fromiter((f.flat.shape[0] for f in flats), dtype=np.int64, count=len(flats))
    var_seg = np.concatenate([[0], np.cumsum(n_per)])
    clause_seg = np.concatenate([[0], np.cumsum(m_per)])
    lens = np.concatenate([f.lens for f in flats])
    flat = np.concatenate([f.flat for f in flats]).astype(np.int64)
    shift = np.repeat(var_seg[:-1], e_per)                  # every literal moves by its formula's variable offset
    flat = np.where(flat < 0, flat - shift, flat + shift)
    return _graph_from_flat(int(var_seg[-1]), lens, flat, var_seg, clause_seg, None, native)


def unit_graph_from_reference_coo(indices, dense_shape, variables_graph=None, clauses_graph=None):
    """Rebuild a unit graph from the reference's ``adj_matrix`` COO ([E,2] literal-row/clause
    pairs, dense shape [2N, M]) and optional per-node graph ids ([N] and [M], ascending).

    This is what ``QuerySAT.call`` uses when it is handed reference-shaped inputs."""
    indices = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
    two_n, m = int(dense_shape[0]), int(dense_shape[1])
    n = two_n // 2
    clauses = [[] for _ in range(m)]
    # keep the storage order inside each clause so duplicates and order survive
    for row, col in indices:
        clauses[col].append(int(row + 1) if row < n else -int(row - n + 1))

    def seg(ids, count):
        if ids is None:
            return [0, count]
        ids = np.asarray(ids, dtype=np.int64)
        if ids.size and np.any(np.diff(ids) < 0):
            raise ValueError("graph ids must be ascending (nodes of one graph contiguous)")
        g = int(ids.max()) + 1 if ids.size else 1
        return np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=g))]).tolist()

    return build_unit_graph(n, clauses, seg(variables_graph, n), seg(clauses_graph, m))
