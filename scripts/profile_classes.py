"""Per-kernel-class device time of the rounds: python scripts/profile_classes.py [chains] [precision code] [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionsat_b200 import _lib, synth, weights, graph
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
prec = int(prec) if prec.isdigit() else _lib.PRECISIONS[prec]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ctx = _lib.Context(0); ctx.set_model(weights.init_weights(seed=1234)); ctx.set_precision(prec)
nv, cl = synth.random_3sat(n, seed=0); ctx.set_graph(graph.build_unit_graph(nv, cl), chains=chains, group_graphs=31)
ctx.profile_rounds(2)
p = ctx.profile_rounds(4)
tot = sum(v[0] for v in p.values())
for k, v in p.items():
    if v[1]: print("%-16s %8.3f ms/round %5.1f%%" % (k, v[0] / 4, 100 * v[0] / tot))
print("total ms/round %.3f  -> est %.0f samples/s" % (tot / 4, chains / (tot / 4 * 1024 / 1e3)))
