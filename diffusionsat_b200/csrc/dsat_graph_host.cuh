// Host side of the graph build (dsat_graph_build): CSR (clause -> literal codes) and CSC (literal code -> clauses) of a
// CNF formula or a disjoint union of formulas from its signed literals laid end to end.  No device work.
//
// Same arrays as diffusionsat_b200/graph.py:_graph_arrays_numpy (which restates reference data/dimac.py:14-18 +
// data/SatSpecifics.py:21-69): inside a clause the literals are ordered by the reference's literal row -- positives by
// variable, then negatives by variable, repeated literals kept -- and the clauses of a literal are ascending with
// repeats kept.  Two stable argsorts over all edges in numpy; here three counting passes without a comparison (sorting each
// clause by insertion was 5x slower: random keys, one mispredicted branch per step):
//   1. degree of every literal code                       -> lit_rowptr
//   2. clauses in order, scatter j to its literals' rows   -> lit_clause (ascending, repeats kept)
//   3. literal rows in reference order (all positives by variable, then all negatives), scatter the code to the rows of
//      its clauses                                         -> cl_lit (every clause ends up ordered by literal row)
#pragma once
#include <stdint.h>

#include <vector>

namespace dsat {

// returns 0, or -(1 + index of the first clause that holds a literal of magnitude 0 or > n_vars)
inline long long graph_build_host(int n_vars, int n_clauses, const int32_t* lens, const int32_t* flat, int32_t* cl_rowptr,
                                  int32_t* cl_lit, int32_t* lit_rowptr, int32_t* lit_clause) {
    const int n_codes = 2 * n_vars;
    for (int l = 0; l <= n_codes; ++l) lit_rowptr[l] = 0;
    cl_rowptr[0] = 0;
    long long e = 0;
    for (int j = 0; j < n_clauses; ++j) {
        const int len = lens[j];
        for (int k = 0; k < len; ++k) {
            const int lit = flat[e + k];
            const int var = (lit < 0 ? -lit : lit) - 1;
            if (lit == 0 || var >= n_vars) return -(1 + (long long)j);
            lit_rowptr[2 * var + (lit < 0) + 1]++;
        }
        e += len;
        cl_rowptr[j + 1] = (int32_t)e;
    }
    for (int l = 0; l < n_codes; ++l) lit_rowptr[l + 1] += lit_rowptr[l];
    std::vector<int32_t> cursor(lit_rowptr, lit_rowptr + n_codes);
    e = 0;
    for (int j = 0; j < n_clauses; ++j) {
        const int len = lens[j];
        for (int k = 0; k < len; ++k) {
            const int lit = flat[e + k];
            lit_clause[cursor[lit < 0 ? 2 * (-lit - 1) + 1 : 2 * (lit - 1)]++] = j;
        }
        e += len;
    }
    cursor.assign(cl_rowptr, cl_rowptr + n_clauses);
    for (int sign = 0; sign < 2; ++sign)
        for (int var = 0; var < n_vars; ++var) {
            const int code = 2 * var + sign;
            for (int t = lit_rowptr[code]; t < lit_rowptr[code + 1]; ++t) cl_lit[cursor[lit_clause[t]]++] = code;
        }
    return 0;
}

}  // namespace dsat
