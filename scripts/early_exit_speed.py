"""Work skipped for finished early-exit groups: trained fixture weights on an easy planted formula with small groups.
python scripts/early_exit_speed.py [precision] [group]   (run with DSAT_SKIP_DONE=0 for the baseline)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diffusionsat_b200 import _lib, graph, synth
from diffusionsat_b200.weights import load_weights
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
group = int(sys.argv[2]) if len(sys.argv) > 2 else 4
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, clauses, _ = synth.planted_3sat(20, 80, seed=1)
ctx = _lib.Context(0)
ctx.set_model(load_weights(os.path.join(root, "tests", "golden", "trained_small.npz")))
ctx.set_precision(prec)
chains = 16384 // group * group
ctx.set_graph(graph.build_unit_graph(n, clauses), chains=chains, group_graphs=group)
ctx.sample_enqueue(32, 32, seed=1); ctx.synchronize()
ctx.timer_begin(); ctx.sample_enqueue(32, 32, seed=2); ms = ctx.timer_end()
packed, sat, latch, _ = ctx.sample_fetch()
print("%s group=%d skip=%s: %.1f ms, %.0f samples/s, sat rate %.3f, checksum %016x" % (
    prec, group, os.environ.get("DSAT_SKIP_DONE", "1"), ms, chains / ms * 1e3, sat.mean(),
    int(np.bitwise_xor.reduce(packed[:, 0] * np.arange(1, chains + 1, dtype=np.uint64)))))
