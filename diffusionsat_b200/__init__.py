"""diffusionsat_b200: B200-native (sm_100a CUDA) sampling hot path of DiffusionSAT.

Public surface mirrors the reference's Python API for the path:
``DiffusionSampler(model_path, dimacs_filename).samples(n)`` and ``QuerySAT.diffusion_step``.
The heavy modules load ``csrc/libdsat.so`` through ctypes on first use; there is no CPU fallback.
"""

from .dimacs import DimacsFile
from .variable_assignment import VariableAssignment
from .graph import UnitGraph, build_unit_graph, build_union_graph, compute_adj_indices
from .weights import QuerySATWeights, init_weights, load_weights, save_weights

__all__ = [
    "DimacsFile", "VariableAssignment", "UnitGraph", "build_unit_graph", "build_union_graph", "compute_adj_indices",
    "QuerySATWeights", "init_weights", "load_weights", "save_weights", "QuerySAT", "DiffusionSampler", "diffusion",
]


def __getattr__(name):  # lazy: these import the CUDA library
    if name == "QuerySAT":
        from .query_sat import QuerySAT
        return QuerySAT
    if name in ("DiffusionSampler", "diffusion"):
        from . import sampler
        return getattr(sampler, name)
    raise AttributeError(name)
