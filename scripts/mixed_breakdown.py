import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from diffusionsat_b200 import _lib, synth, weights, graph as G, dist as D
rng = np.random.default_rng(0)
formulas = []
for i in range(1200):
    nv = int(rng.integers(3, 101)); formulas.append(synth.random_ksat_mixed(nv, max(1, int(4.3 * nv)), seed=1000 + i))
formulas = [G.flatten_formula(*f) for f in formulas]      # as bench.py --config mixed does (loader side, once)
ctx = _lib.Context(0); ctx.set_model(weights.init_weights(seed=1234)); ctx.set_precision("fp32")
batches = D.pack_batches(formulas)
t = {"union": 0.0, "set_graph": 0.0, "model_call": 0.0}
for rep in range(2):
    for k in t: t[k] = 0.0
    for b, idxs in enumerate(batches):
        t0 = time.perf_counter(); unit = G.build_union_graph([formulas[i] for i in idxs]); t1 = time.perf_counter()
        ctx.set_graph(unit, chains=1, group_graphs=0); t2 = time.perf_counter()
        bits = np.random.default_rng([0, b]).integers(0, 2, unit.n_vars).astype(np.float32)
        pred, steps, _ = ctx.model_call(0.5, np.stack([bits, 1 - bits], axis=1), rounds=32, seed=b); t3 = time.perf_counter()
        t["union"] += t1 - t0; t["set_graph"] += t2 - t1; t["model_call"] += t3 - t2
    print("rep %d: %d batches, per batch ms: %s" % (rep, len(batches), {k: round(1e3 * v / len(batches), 1) for k, v in t.items()}))
t0 = time.perf_counter(); D.forward_formulas_sharded(lambda r: ctx, formulas, 0.5, rounds=32, seed=0); t1 = time.perf_counter()
print("forward_formulas_sharded (graph build on the helper thread): %.1f ms per batch, %.0f formulas/s" % (1e3 * (t1 - t0) / len(batches), len(formulas) / (t1 - t0)))
