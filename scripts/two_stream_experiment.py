"""Do two half-size sampling runs on two streams overlap (tensor-bound MLP kernels of one with HBM-bound gathers of the
other)?  python scripts/two_stream_experiment.py [chains] [precision] [rounds]
Prints wall/device time of one context with `chains` chains against two contexts with chains/2 each, driven by two threads."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionsat_b200 import _lib, build, graph, synth, weights

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
build.build()
nv, clauses = synth.random_3sat(100, seed=0)
unit = graph.build_unit_graph(nv, clauses)
wts = weights.init_weights(seed=1234)


def make(c):
    ctx = _lib.Context(0)
    ctx.set_model(wts)
    ctx.set_precision(_lib.PRECISIONS[prec])
    ctx.set_graph(unit, chains=c, group_graphs=31)
    ctx.sample_enqueue(1, 2, seed=1)
    ctx.synchronize()
    return ctx


def run(ctxs, steps=2):
    def work(c, i):
        c.sample_enqueue(steps, rounds, seed=5 + i, chain_offset=i * c.chains)
        c.synchronize()
    ts = [threading.Thread(target=work, args=(c, i)) for i, c in enumerate(ctxs)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


one = make(chains)
run([one]); w1 = run([one])
one.close()
two = [make(chains // 2), make(chains // 2)]
run(two); w2 = run(two)
per_round = lambda w: 1e3 * w / (2 * rounds)
print("%s, %d chains, %d rounds/step: one stream %.3f ms/round, two streams %.3f ms/round (x%.2f)" %
      (prec, chains, rounds, per_round(w1), per_round(w2), w1 / w2))
