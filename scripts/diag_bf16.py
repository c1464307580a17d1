import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diffusionsat_b200 import _lib, graph as G, synth
from oracle import querysat_oracle as O
from tests import helpers as H
n_vars, chains, seed = 30, 5, 0
_, clauses = synth.random_3sat(n_vars, seed=seed)
wts = H.make_weights(seed=11)
ctx = _lib.Context(0); ctx.set_model(wts); ctx.set_precision(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ctx.set_graph(G.build_unit_graph(n_vars, clauses), chains=chains, group_graphs=0)
n_rows, rounds = n_vars * chains, 2
noise = H.noise_for(n_rows, rounds, seed)
noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
graph, _, trace = H.oracle_trace(n_vars, clauses, chains, wts, 0.625, noisy, noise, rounds)
ctx.debug_begin(0.625, noisy, noise["labels"])
F = Q = 128; AUX = 16
def per_chain(name, got, want, rows_per_chain):
    errs = []
    for c in range(chains):
        s = slice(c * rows_per_chain, (c + 1) * rows_per_chain)
        errs.append(np.abs(got[s] - want[s]).max() / (np.abs(want).max() + 1e-9))
    print("%-14s" % name, " ".join("%.1e" % e for e in errs))
for r in range(rounds):
    ctx.debug_round(r, noise["normals"][r])
    tr = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in trace[r].items()}
    n, m = graph.n_vars, graph.n_clauses
    mc = m // chains
    vrow, crow = ctx.debug_read("VROW"), ctx.debug_read("CROW")
    print("round", r)
    per_chain("query", ctx.debug_read("QS")[:, :Q], tr["query"], n_vars)
    per_chain("lit", ctx.debug_read("LIT"), tr["var_msg"], n_vars)
    per_chain("cmsg", crow[:, F:F + Q], tr["clause_messages"], mc)
    per_chain("cl4", crow[:, F + Q:], 4 * tr["clauses_loss"], mc)
    per_chain("clause_data", ctx.debug_read("COUT"), tr["clause_data"], mc)
    cs = tr["clause_state"]; per_chain("clause_state", crow[:, :F], cs * np.float32(0.2) + cs * np.float32(0.8), mc)
    per_chain("grad", vrow[:, F + AUX:F + AUX + Q], tr["variables_grad"], n_vars)
    per_chain("loss_pos", vrow[:, F + AUX + Q:F + AUX + 2 * Q], tr["variables_loss"][:n], n_vars)
    per_chain("loss_neg", vrow[:, F + AUX + 2 * Q:], tr["variables_loss"][n:], n_vars)
    per_chain("variables", ctx.debug_read("SPRE"), tr["variables"], n_vars)
    per_chain("logits", ctx.debug_read("LOGITS")[:, :8], tr["logits"], n_vars)
