"""Truth assignment with the reference's integer encoding (host side).

Mirror of ``utils/VariableAssignment.py`` (reference):

* ``__int__``                   -> ``:63-69``  bit i of the int = variable i+1 (x1 is the LSB)
* ``assign_all_from_bit_list``  -> ``:52-57``
* ``assign_all_from_int``       -> ``:59-61``
* ``satisfiable``               -> ``:79-90``
* vector length when only clauses are given = max |literal| -> ``:34-36``

The GPU path produces the same encoding as packed 64-bit words
(``csrc/dsat_kernels.cu: pack_assignments_kernel``); this class is the host
type the sampler API hands out and the checker the tests use.
"""

from __future__ import annotations

import numpy as np


class VariableAssignment:
    def __init__(self, n_vars=0, clauses=[]):
        if clauses == [] and type(n_vars) != int:
            # positional clauses, as the reference tolerates
            clauses, n_vars = n_vars, 0
        if type(clauses) != list:
            clauses = [[int(v) for v in row] for row in clauses]
        if n_vars == 0:
            flat = np.array([lit for clause in clauses for lit in clause])
            n_vars = max(abs(flat.max()), abs(flat.min()))
        self.x = [False] * int(n_vars)
        self.clauses = clauses

    def assign(self, i: int, value: bool):
        self.x[i] = value  # IndexError when i is past the vector, as in the reference

    def assign_all(self, x):
        self.x = x

    def assign_all_from_int_list(self, x):
        for lit in x:
            self.assign(abs(lit) - 1, lit > 0)

    def assign_all_from_bit_list(self, x):
        for pos, bit in enumerate(x):
            self.assign(pos, int(bit) == 1)

    def assign_all_from_int(self, i):
        for pos in range(len(self.x)):
            self.assign(pos, (i >> pos) & 1 == 1)

    def __int__(self):
        value = 0
        for pos, bit in enumerate(self.x):
            if bit:
                value |= 1 << pos
        return value

    def __str__(self):
        return "".join("1" if bit else "0" for bit in self.x)

    def satisfiable(self):
        for clause in self.clauses:
            if not any((lit > 0) == self.x[abs(lit) - 1] for lit in clause):
                return False
        return True

    def as_int_list(self):
        return [(pos + 1) if bit else -(pos + 1) for pos, bit in enumerate(self.x)]

    def value(self, i):
        return self.x[i]

    def values(self):
        return self.x
