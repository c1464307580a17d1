"""A/B of the standalone segment-sum SpMM instantiations (register budget x descriptor prefetch) on BASELINE configs[4]'s
graph (n=10000, m=43000): every variant is first checked against a host segment sum on a small mixed k-SAT graph
(fp32 bit-exact: same summation order), then timed over the feature width x storage type x direction sweep.

  python scripts/spmm_variants.py [--variants 8:0,6:0,5:0,4:0,5:1,4:1] [--only F128_bf16_0] [--no-check] [--out file]

A variant is MINB:PF[:LIT_DW] (DSAT_SPMM_MINB, DSAT_SPMM_PF, DSAT_SPMM_LIT_DW; read when the context is created).
"default" = the built-in table (`spmm_plan` in dsat_api.cu); "half" = DSAT_SPMM_HALF=1 (half the lanes per row)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from diffusionsat_b200 import _lib, graph, synth, weights

SHAPES = ((128, "f32"), (64, "f32"), (256, "f32"), (128, "bf16"), (64, "bf16"), (256, "bf16"))
PEAK = 6552.3


def make_context(variant, wts):
    for key in ("DSAT_SPMM_MINB", "DSAT_SPMM_PF", "DSAT_SPMM_LIT_DW", "DSAT_SPMM_HALF"):
        os.environ.pop(key, None)
    if variant == "half":
        os.environ["DSAT_SPMM_HALF"] = "1"
    elif variant != "default":
        parts = variant.split(":")
        os.environ["DSAT_SPMM_MINB"], os.environ["DSAT_SPMM_PF"] = parts[0], parts[1]
        if len(parts) > 2:
            os.environ["DSAT_SPMM_LIT_DW"] = parts[2]
    ctx = _lib.Context(0)
    ctx.set_model(wts)
    return ctx


def check(ctx, dev):
    """Small mixed k-SAT graph, 3 chains: rows of 1..8 entries, several rows per warp pass, ragged last pass."""
    n_vars, chains = 50, 3
    _, clauses = synth.random_ksat_mixed(n_vars, 180, seed=4)
    unit = graph.build_unit_graph(n_vars, clauses)
    ctx.set_graph(unit, chains=chains, group_graphs=0)
    gen = torch.Generator().manual_seed(0)
    for feat, dt in SHAPES:
        tdt = torch.float32 if dt == "f32" else torch.bfloat16
        for direction, rows_in, rows_out, rowptr, col, scale in (
                (0, 2 * n_vars, unit.n_clauses, unit.cl_rowptr, unit.cl_lit, unit.rev_degree_weight()),
                (1, unit.n_clauses, 2 * n_vars, unit.lit_rowptr, unit.lit_clause, unit.degree_weight())):
            x = torch.randn(chains, rows_in, feat, generator=gen).to(tdt)
            xd = x.to(dev)
            yd = torch.zeros(chains, rows_out, feat, dtype=tdt, device=dev)
            torch.cuda.synchronize()
            ctx.spmm(direction, xd.data_ptr(), yd.data_ptr(), feat, 0 if dt == "f32" else 1, chains)
            ctx.synchronize()
            xf = x.float().numpy()
            want = np.zeros((chains, rows_out, feat), dtype=np.float32)
            for r in range(rows_out):
                acc = np.zeros((chains, feat), dtype=np.float32)
                for e in range(rowptr[r], rowptr[r + 1]):
                    acc = acc + xf[:, col[e]]
                want[:, r] = acc * scale[r]
            got = yd.float().cpu().numpy()
            if dt == "f32":
                if not np.array_equal(got, want):
                    return "MISMATCH F=%d %s dir=%d" % (feat, dt, direction)
            elif not np.allclose(got, want, rtol=1e-2, atol=1e-2):
                return "MISMATCH F=%d %s dir=%d" % (feat, dt, direction)
    return "ok"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="default")
    ap.add_argument("--only", default="", help="e.g. F128_bf16_0 (direction 0 = clause<-literal)")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    out = open(args.out, "a") if args.out else None

    def say(line):
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    dev = torch.device("cuda:0")
    wts = weights.init_weights(seed=1)
    variants = args.variants.split(",")
    ctxs = {}
    for v in variants:
        ctxs[v] = make_context(v, wts)
        if not args.no_check:
            say("# variant %s: small-graph check %s" % (v, check(ctxs[v], dev)))
    n, m = 10000, 43000
    nv, cl = synth.random_3sat(n, m, seed=5)
    unit = graph.build_unit_graph(nv, cl)
    for v in variants:
        ctxs[v].set_graph(unit, chains=1, group_graphs=0)
    for feat, dt in SHAPES:
        tdt, code, es = (torch.float32, 0, 4) if dt == "f32" else (torch.bfloat16, 1, 2)
        chains = max(8, int(3.0e9 / ((2 * n + m) * feat * es)))
        for name, d, rin, rout in (("clause<-literal", 0, 2 * n, m), ("literal<-clause", 1, m, 2 * n)):
            if args.only and args.only != "F%d_%s_%d" % (feat, dt, d):
                continue
            x = torch.randn(chains, rin, feat, device=dev).to(tdt)
            y = torch.empty(chains, rout, feat, device=dev, dtype=tdt)
            torch.cuda.synchronize()
            nbytes = (rin + rout) * chains * feat * es + (unit.nnz + rout + 1) * 4
            cells = []
            for v in variants:
                ctx = ctxs[v]
                for _ in range(3):
                    ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
                ctx.synchronize()
                ctx.timer_begin()
                for _ in range(args.reps):
                    ctx.spmm(d, x.data_ptr(), y.data_ptr(), feat, code, chains)
                ms = ctx.timer_end() / args.reps
                cells.append("%s %.3f ms %.0f GB/s (%.0f %%)" % (v, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / PEAK * 100))
            say("F=%d %s %s chains=%d: %s" % (feat, dt, name, chains, " | ".join(cells)))
            del x, y


if __name__ == "__main__":
    main()
