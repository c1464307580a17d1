"""Quick device timing of the sampling path (not the bench): python scripts/quick_time.py [chains] [steps] [rounds] [precision]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diffusionsat_b200 import _lib, build, synth, weights, graph

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
n = int(sys.argv[5]) if len(sys.argv) > 5 else 100
build.build()
ctx = _lib.Context(0)
ctx.set_model(weights.init_weights(seed=1234))
ctx.set_precision(_lib.PRECISIONS[prec])
nv, clauses = synth.random_3sat(n, seed=0)
unit = graph.build_unit_graph(nv, clauses)
ctx.set_graph(unit, chains=chains, group_graphs=graph.chains_per_reference_batch(nv, len(clauses)))
ctx.sample_enqueue(1, 2, seed=1); ctx.synchronize()
l0 = ctx.launch_count()
ctx.timer_begin()
ctx.sample_enqueue(steps, rounds, seed=2)
ms = ctx.timer_end()
l1 = ctx.launch_count()
cr = chains * steps * rounds
flop = 2 * (723483 * n + 130560 * len(clauses)) * cr
print("chains=%d n=%d m=%d steps=%d rounds=%d prec=%s: %.2f ms, %.3f us/chain-round, %.1f TFLOP/s, est %.1f samples/s at 32x32, launches=%d"
      % (chains, n, len(clauses), steps, rounds, prec, ms, 1e3 * ms / cr, flop / ms / 1e9, chains / (ms / 1e3 * 1024 / (steps * rounds)), l1 - l0))
packed, is_sat, latch, _ = ctx.sample_fetch()
print("sat rate", is_sat.mean(), "first ints", [hex(int(x)) for x in packed[:3, 0]])
