"""Standalone segment-sum SpMM sweep on BASELINE configs[4]'s graph (n=10000, m=43000): feature width x storage type x
direction, compulsory bytes / CUDA-event time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionsat_b200 import _lib, synth, weights, graph
n, m = 10000, 43000
ctx = _lib.Context(0); ctx.set_model(weights.init_weights(seed=1))
nv, cl = synth.random_3sat(n, m, seed=5); unit = graph.build_unit_graph(nv, cl)
ctx.set_graph(unit, chains=1, group_graphs=0)
dev = torch.device("cuda:0")
for feat, tdt, code, es in ((128, torch.float32, 0, 4), (64, torch.float32, 0, 4), (256, torch.float32, 0, 4), (128, torch.bfloat16, 1, 2),
                            (64, torch.bfloat16, 1, 2), (256, torch.bfloat16, 1, 2)):
    chains = max(8, int(3.0e9 / ((2 * n + m) * feat * es)))
    for name, d, rin, rout in (("clause<-literal", 0, 2 * n, m), ("literal<-clause", 1, m, 2 * n)):
        x = torch.randn(chains, rin, feat, device=dev).to(tdt); y = torch.empty(chains, rout, feat, device=dev, dtype=tdt)
        torch.cuda.synchronize()
        for _ in range(3): ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
        ctx.synchronize(); ctx.timer_begin()
        for _ in range(5): ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
        ms = ctx.timer_end() / 5
        nbytes = (rin + rout) * chains * feat * es + (unit.nnz + rout + 1) * 4
        print("F=%d %s %s chains=%d: %.3f ms  %.0f GB/s  (%.0f %% of 6552)" % (feat, str(tdt)[6:], name, chains, ms,
              nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.52))
        del x, y
