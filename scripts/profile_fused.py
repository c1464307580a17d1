import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionsat_b200 import _lib, synth, weights, graph
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _lib.Context(0); ctx.set_model(weights.init_weights(seed=1234)); ctx.set_precision(1)
nv, cl = synth.random_3sat(100, seed=0); ctx.set_graph(graph.build_unit_graph(nv, cl), chains=chains, group_graphs=31)
ctx.profile_rounds(2)
names = ["query", "literal", "clause", "update", "output"]
for w in range(5):
    c = ctx.profile_fused(w)
    tot = c[0] or 1
    tiles = {0: chains * 100, 1: chains * 100, 2: chains * len(cl), 3: chains * 100, 4: chains * 100}[w] / 128 / 148
    print("%-8s total %8d cyc (%.0f cyc/tile) | producer: ah_free %4.1f%% ring_empty %4.1f%% | mma: tmem_empty %4.1f%% a_full %4.1f%% h_full %4.1f%% ring_full %4.1f%% | epi: tmem_full %4.1f%% hidden %4.1f%% final %4.1f%%"
          % (names[w], tot, tot / tiles, 100 * c[1] / tot, 100 * c[2] / tot, 100 * c[3] / tot, 100 * c[4] / tot, 100 * c[5] / tot,
             100 * c[6] / tot, 100 * c[7] / tot, 100 * c[8] / tot, 100 * c[9] / tot), "| mma issue %4.1f%% commit %4.1f%% | mma k-loop total %4.1f%% tile commit %4.1f%%" % (100 * c[12] / tot, 100 * c[13] / tot, 100 * c[14] / tot, 100 * c[15] / tot))
