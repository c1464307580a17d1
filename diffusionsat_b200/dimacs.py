"""DIMACS CNF reader/writer with the reference's parsing rules.

Host-side mirror of ``utils/DimacsFile.py`` (reference). The method names,
argument meaning and error behaviour follow the reference class so that code
written against it keeps working:

* ``load_from_lines``  -> reference ``utils/DimacsFile.py:49-85``
* ``add_clauses``      -> reference ``utils/DimacsFile.py:102-108``
* ``reduce_clauses``   -> reference ``utils/DimacsFile.py:110-128`` (+ ``:130-142``)
* ``is_satisfiable``   -> reference ``utils/DimacsFile.py:163-180``
* ``store``/``__str__``-> reference ``utils/DimacsFile.py:182-218``

Parsing rules that must be matched bit-exactly:
one clause per line; tokens after the first ``0`` are ignored; a line that
holds only ``0`` adds an empty clause; a line containing ``p cnf`` sets
``n_vars`` from the first integer after it; lines whose first character is a
letter are ignored unless they start with ``v`` (assignment lines); lines
starting with ``--`` are ignored; anything else that is not an integer (for
example the SATLIB ``%`` trailer) raises ``ValueError``.
"""

from __future__ import annotations

import numpy as np


def _as_clause_lists(clauses):
    """Accept nested sequences / ragged arrays and return ``list[list[int]]``."""
    if isinstance(clauses, list):
        return clauses
    return [[int(v) for v in row] for row in clauses]


def _max_abs_literal(clauses) -> int:
    flat = np.array([lit for clause in clauses for lit in clause])
    return int(max(abs(flat.max()), abs(flat.min())))


def _contained_in(small, big) -> bool:
    """True when every literal of sorted ``small`` also occurs in sorted ``big``."""
    j = 0
    for lit in small:
        while j < len(big) and big[j] < lit:
            j += 1
        if j >= len(big) or big[j] != lit:
            return False
        j += 1
    return True


class DimacsFile:
    def __init__(self, filename="file.cnf", n_vars=0, clauses=[]):
        clauses = _as_clause_lists(clauses)
        if n_vars == 0 and len(clauses) > 0:
            n_vars = _max_abs_literal(clauses)
        self.filename = filename
        self.n_vars = n_vars
        self.i_clauses = clauses
        self.b_values = {}
        self.comments = []

    # ------------------------------------------------------------------ load
    def load(self):
        with open(self.filename, "r") as handle:
            self.load_from_lines(handle.readlines())

    def load_from_string(self, text):
        self.load_from_lines([ln.strip() for ln in text.split("\n")])

    def load_from_lines(self, lines):
        self.n_vars = 0
        self.i_clauses = []
        for raw in lines:
            line = raw.strip()
            if not line:
                continue
            if "p cnf" in line:
                # the reference drops the first five characters of the line
                # (not of the match) and reads up to the next blank
                rest = line[len("p cnf"):].strip()
                self.n_vars = int(rest[:rest.find(" ")])
                continue
            head = line[0]
            if head.isalpha():
                if head == "v":
                    for tok in line[1:].strip().split():
                        lit = int(tok)
                        if lit > 0:
                            self.b_values[lit] = True
                        if lit < 0:
                            self.b_values[-lit] = False
                continue
            if line.startswith("--"):
                continue
            clause = []
            for tok in line.split():
                lit = int(tok)  # ValueError for '%' and other junk, as in the reference
                if lit == 0:
                    break
                clause.append(lit)
            self.add_clause(clause)

    # --------------------------------------------------------------- queries
    def number_of_vars(self):
        return self.n_vars

    def number_of_clauses(self):
        return len(self.i_clauses)

    def clauses(self):
        return self.i_clauses

    # -------------------------------------------------------------- mutation
    def add_comment(self, comment):
        self.comments.append(comment)

    def add_clause(self, clause):
        self.add_clauses([clause])

    def add_clauses(self, clauses):
        for clause in clauses:
            for lit in clause:
                self.n_vars = max(self.n_vars, abs(lit))
            self.i_clauses.append(clause)

    def reduce_clauses(self):
        """Drop duplicate clauses, then clauses subsumed by a shorter one."""
        unique = [list(c) for c in {tuple(sorted(c)) for c in self.i_clauses}]
        unique.sort(key=len)
        kept = []
        for clause in unique:
            if not any(_contained_in(k, clause) for k in kept):
                kept.append(clause)
        self.i_clauses = kept

    # ------------------------------------------------------------ assignment
    def set_value(self, i, value):
        self.b_values[abs(i)] = value

    def set_values(self, dict_of_values):
        for key, value in dict_of_values.items():
            self.set_value(key, value)

    def get_value(self, i):
        return self.b_values[abs(i)]

    def is_satisfiable(self):
        for var in range(1, self.n_vars + 1):
            if var not in self.b_values:
                raise Exception(
                    "Not all variables have values. Variable " + str(var) + " does not."
                )
        for clause in self.clauses():
            if not any((lit > 0) == bool(self.get_value(lit)) for lit in clause if lit != 0):
                return False
        return True

    # ------------------------------------------------------------------ dump
    def __str__(self):
        out = ["p cnf %d %d" % (self.number_of_vars(), self.number_of_clauses())]
        for clause in self.clauses():
            out.append("".join(str(lit) + " " for lit in clause) + "0")
        return "\n".join(out) + "\n"

    def store(self, *comments):
        with open(self.filename, "w") as handle:
            for c in list(self.comments) + list(comments):
                handle.write("c " + c + "\n")
            handle.write(str(self))
