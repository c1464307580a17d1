import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionsat_b200 import _lib, synth, weights, graph
n, m, feat = 10000, 43000, 128
ctx = _lib.Context(0); ctx.set_model(weights.init_weights(seed=1))
nv, cl = synth.random_3sat(n, m, seed=5); unit = graph.build_unit_graph(nv, cl)
ctx.set_graph(unit, chains=1, group_graphs=0)
dev = torch.device("cuda:0")
for tdt, code, es in ((torch.float32, 0, 4), (torch.bfloat16, 1, 2)):
    chains = int(3.0e9 / ((2 * n + m) * feat * es))
    for name, d, rin, rout in (("clause<-literal", 0, 2 * n, m), ("literal<-clause", 1, m, 2 * n)):
        x = torch.randn(chains, rin, feat, device=dev).to(tdt); y = torch.empty(chains, rout, feat, device=dev, dtype=tdt)
        torch.cuda.synchronize()
        for _ in range(3): ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
        ctx.synchronize(); ctx.timer_begin()
        for _ in range(5): ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
        ms = ctx.timer_end() / 5
        nbytes = (rin + rout) * chains * feat * es + (unit.nnz + rout + 1) * 4
        print("%s %s chains=%d: %.3f ms  %.0f GB/s" % (name, str(tdt)[6:], chains, ms, nbytes / ms / 1e6))
        del x, y
