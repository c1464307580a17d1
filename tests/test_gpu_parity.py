"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs and injected noise.  Tolerances: integer/bit/index results exact; fp32 activations within
1e-4 relative per teacher-forced round and 1e-3 relative for free-running logits (BASELINE.json);
discrete decisions are compared away from ties only."""
import numpy as np
import pytest
import torch

from diffusionsat_b200 import graph as G
from diffusionsat_b200 import philox, synth
from oracle import querysat_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

F = Q = 128
AUX = 16


def carry(x):
    x = np.asarray(x, dtype=np.float32)
    return (x * np.float32(0.2) + x * np.float32(0.8)).astype(np.float32)


def bind(ctx, n_vars, clauses, chains, wts, group=0):
    unit = G.build_unit_graph(n_vars, clauses)
    ctx.set_model(wts)
    ctx.set_graph(unit, chains=chains, group_graphs=group)
    return unit


def check(name, got, want, tol):
    err = H.rel_err(got, want)
    assert err < tol, "%s: relative error %.3e >= %.1e" % (name, err, tol)


@pytest.mark.parametrize("n_vars,chains,seed", [(30, 5, 0), (12, 3, 1)])
def test_round_teacher_forced(ctx, n_vars, chains, seed):
    _, clauses = synth.random_3sat(n_vars, seed=seed)
    wts = H.make_weights(seed=11 + seed)
    bind(ctx, n_vars, clauses, chains, wts)
    n_rows, rounds = n_vars * chains, 4
    noise = H.noise_for(n_rows, rounds, seed)
    noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
    noise_scale = 0.625
    graph, _, trace = H.oracle_trace(n_vars, clauses, chains, wts, noise_scale, noisy, noise, rounds)
    ctx.debug_begin(noise_scale, noisy, noise["labels"])
    n, m = graph.n_vars, graph.n_clauses
    for r in range(rounds):
        tr = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in trace[r].items()}
        if r > 0:   # teacher forcing: start the round from the oracle's carried state
            prev = trace[r - 1]
            vrow = ctx.debug_read("VROW"); vrow[:, :F] = carry(prev["variables"].numpy()); ctx.debug_write("VROW", vrow)
            crow = ctx.debug_read("CROW"); crow[:, :F] = carry(prev["clause_state"].numpy()); ctx.debug_write("CROW", crow)
        ctx.debug_round(r, noise["normals"][r])
        vrow, crow = ctx.debug_read("VROW"), ctx.debug_read("CROW")
        qs, lit, cout = ctx.debug_read("QS"), ctx.debug_read("LIT"), ctx.debug_read("COUT")
        tol = 1e-4
        # aux columns are copies (exact on the CUDA-core path; hi + lo bf16 planes on the default tensor-core path)
        np.testing.assert_allclose(vrow[:, F:F + 9], tr["v1"][:, F:F + 9], rtol=2.0 ** -16, atol=0)
        check("query", qs[:, :Q], tr["query"], tol)
        check("softplus(+q)", qs[:, Q:2 * Q], np.logaddexp(0, tr["query"].astype(np.float64)), tol)
        check("softplus(-q)", qs[:, 2 * Q:], np.logaddexp(0, -tr["query"].astype(np.float64)), tol)
        check("lit_query", lit, tr["var_msg"], tol)
        check("clause_messages", crow[:, F:F + Q], tr["clause_messages"], tol)
        check("4*clauses_loss", crow[:, F + Q:], 4 * tr["clauses_loss"], tol)
        check("clause_data", cout, tr["clause_data"], tol)
        check("clause_state", crow[:, :F], carry(tr["clause_state"]), tol)
        check("variables_grad", vrow[:, F + AUX:F + AUX + Q], tr["variables_grad"], tol)
        check("loss_pos", vrow[:, F + AUX + Q:F + AUX + 2 * Q], tr["variables_loss"][:n], tol)
        check("loss_neg", vrow[:, F + AUX + 2 * Q:], tr["variables_loss"][n:], tol)
        check("variables", ctx.debug_read("SPRE"), tr["variables"], tol)
        check("variables carried", vrow[:, :F], carry(tr["variables"]), tol)
        logits = ctx.debug_read("LOGITS")
        check("logits", logits[:, :8], tr["logits"], 5e-4)
        assert np.all(logits[:, 8:] == 0)
        # logit-map choice: compare where the oracle's best and second-best losses are not tied
        pgl = tr["per_graph_loss"]
        srt = np.sort(pgl, axis=1)
        clear = (srt[:, 1] - srt[:, 0]) > 1e-4 * (np.abs(srt[:, 0]) + 1e-6)
        groups = ctx.debug_groups()
        np.testing.assert_array_equal(groups["graph_map"][clear], tr["best_graph_map"][clear])
        same = np.repeat(groups["graph_map"] == tr["best_graph_map"], n_vars)
        out = ctx.debug_read("OUT")[:, 0]
        check("out_logits", out[same], tr["out_logits"][same], 5e-4)


def test_model_call_free_running(ctx):
    n_vars, chains, rounds = 30, 4, 32
    _, clauses = synth.random_3sat(n_vars, seed=3)
    wts = H.make_weights(seed=5)
    bind(ctx, n_vars, clauses, chains, wts)
    n_rows = n_vars * chains
    noise = H.noise_for(n_rows, rounds, 9)
    noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
    graph, out, trace = H.oracle_trace(n_vars, clauses, chains, wts, 0.5, noisy, noise, rounds, dtype=torch.float64)
    pred, steps, loss = ctx.model_call(0.5, noisy, labels=noise["labels"], normals=noise["normals"], rounds=rounds)
    groups = ctx.debug_groups()
    same = np.repeat(groups["graph_map"] == trace[-1]["best_graph_map"].numpy(), n_vars)
    assert same.mean() > 0.5
    check("prediction logits (fp32 vs fp64 oracle, 32 rounds)", pred[same], out[0].numpy()[same], 1e-3)
    assert steps[0] == out[1]
    assert abs(loss[0] - float(out[2])) < 1e-3 * max(1.0, abs(float(out[2])))


def test_device_philox_matches_host_spec(ctx):
    n_vars, chains = 20, 3
    _, clauses = synth.random_3sat(n_vars, seed=2)
    bind(ctx, n_vars, clauses, chains, H.make_weights(seed=1))
    n_rows = n_vars * chains
    noisy = np.tile(np.array([[1.0, 0.0]], dtype=np.float32), (n_rows, 1))
    ctx.debug_begin(0.25, noisy, None)       # labels from Philox, seed 0, step 0
    ctx.debug_round(2, None)                 # normals from Philox, round 2
    vrow = ctx.debug_read("VROW")
    want = philox.normals(0, np.arange(n_rows), 0, 2)
    np.testing.assert_allclose(vrow[:, F:F + 4], want, rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(vrow[:, F + 4:F + 7], np.tile(np.array([1.0, 0.0, 0.25], np.float32), (n_rows, 1)))


def _oracle_diffusion(n_vars, clauses, chains, wts, noise, steps, rounds):
    graph = O.OracleGraph.copies(n_vars, clauses, chains)
    w = O.weights_to_torch(wts, torch.float32)
    trace = []
    acc, final, latch = O.diffusion(steps, graph, w, torch.from_numpy(noise["uniforms"]),
                                    torch.from_numpy(noise["labels"].astype(np.int64)), torch.from_numpy(noise["normals"]),
                                    rounds, trace=trace)
    return graph, acc, final, latch, trace


@pytest.mark.parametrize("n_vars,n_clauses,chains,steps,rounds,seed", [(8, 16, 6, 6, 3, 0), (40, 120, 4, 5, 4, 1)])
def test_sample_matches_oracle_diffusion(ctx, n_vars, n_clauses, chains, steps, rounds, seed):
    _, clauses, _ = synth.planted_3sat(n_vars, n_clauses, seed=seed)
    wts = H.make_weights(seed=21 + seed)
    bind(ctx, n_vars, clauses, chains, wts, group=chains)
    n_rows = n_vars * chains
    noise = H.noise_for(n_rows, rounds, 100 + seed, steps=steps)
    graph, acc, final, latch, trace = _oracle_diffusion(n_vars, clauses, chains, wts, noise, steps, rounds)
    packed, is_sat, latch_step, sat_any = ctx.sample(steps, rounds, uniforms=noise["uniforms"], labels=noise["labels"],
                                                     normals=noise["normals"])
    # decisions can only differ where a probability sits on a rounding boundary: require the oracle's
    # predictions to be clear of 0.5 and of the uniform thresholds, else skip the chain
    from diffusionsat_b200.sampler import unpack_assignments
    got = unpack_assignments(packed, n_vars)
    checked = 0
    for c in range(chains):
        bits = final[c * n_vars:(c + 1) * n_vars]
        want = O.encode_assignment(bits)
        margins = [np.abs(t["predictions"].numpy()[c * n_vars:(c + 1) * n_vars] - 0.5).min() for t in trace]
        if min(margins) < 1e-4:      # 100 x the fp32 paths' measured probability error; with 1e-3 only 2 of 6 chains were checked
            continue
        checked += 1
        assert got[c] == want, "chain %d: %x != %x" % (c, got[c], want)
        assert latch_step[c] == latch[c * n_vars]
        assert bool(is_sat[c]) == O._satisfiable_py([bool(b) for b in bits], clauses)
        assert bool(sat_any[c]) == (latch[c * n_vars] >= 0)
    assert checked >= (chains + 1) // 2, "only %d of %d chains were clear of rounding boundaries" % (checked, chains)


def test_early_exit_is_per_group(ctx):
    """Each group of graphs stops at the first round in which all of ITS graphs are satisfied
    (reference model/query_sat.py:330-338 applied per reference batch)."""
    n_vars, clauses = 3, [[1, 2], [-1, 3], [2, 3]]
    wts = H.make_weights(seed=4)
    chains, group, rounds = 6, 2, 6
    bind(ctx, n_vars, clauses, chains, wts, group=group)
    n_rows = n_vars * chains
    noise = H.noise_for(n_rows, rounds, 17)
    noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
    pred, steps, loss = ctx.model_call(0.4, noisy, labels=noise["labels"], normals=noise["normals"], rounds=rounds)
    w = O.weights_to_torch(wts)
    for gidx in range(chains // group):
        rows = slice(gidx * group * n_vars, (gidx + 1) * group * n_vars)
        graph = O.OracleGraph.copies(n_vars, clauses, group)
        out = O.model_loop(graph, w, 0.4, torch.from_numpy(noisy[rows]), torch.from_numpy(noise["labels"][rows].astype(np.int64)),
                           torch.from_numpy(noise["normals"][:, rows]), rounds)
        assert steps[gidx] == out[1]
        check("group %d prediction" % gidx, pred[rows], out[0].numpy(), 1e-3)
        assert abs(loss[gidx] - float(out[2])) < 1e-3 * max(1.0, abs(float(out[2])))


def test_edge_cases_empty_clause_duplicates_isolated_variable(ctx):
    # variable 5 occurs nowhere (degree 0), clause 2 repeats a literal, clause 3 is a unit clause
    n_vars, clauses = 5, [[1, -2, 3], [2, 2, -4], [4], [-1, -3, 4, 2]]
    wts = H.make_weights(seed=8)
    chains, rounds = 2, 3
    bind(ctx, n_vars, clauses, chains, wts)
    n_rows = n_vars * chains
    noise = H.noise_for(n_rows, rounds, 5)
    noisy = O.randomized_rounding(torch.full((n_rows, 2), 0.5), torch.from_numpy(noise["uniform"])).numpy()
    graph, out, trace = H.oracle_trace(n_vars, clauses, chains, wts, 0.9, noisy, noise, rounds)
    pred, steps, _ = ctx.model_call(0.9, noisy, labels=noise["labels"], normals=noise["normals"], rounds=rounds)
    groups = ctx.debug_groups()
    assert steps[0] == out[1]
    same = np.repeat(groups["graph_map"] == trace[-1]["best_graph_map"].numpy(), n_vars)
    assert same.any(), "the logit-map choice differs from the oracle's for every graph"
    check("prediction", pred[same], out[0].numpy()[same], 1e-3)
    # an empty clause can never be satisfied (reference VariableAssignment.satisfiable / is_batch_sat)
    clauses2 = [[1, 2], [], [-1]]
    bind(ctx, 2, clauses2, 3, wts)
    packed, is_sat, latch_step, _ = ctx.sample(3, 2, seed=1)
    assert not is_sat.any() and (latch_step == -1).all()


@pytest.mark.parametrize("feat", [64, 128, 256])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_spmm_segment_sum(ctx, feat, dtype):
    n_vars, chains = 50, 3
    _, clauses = synth.random_ksat_mixed(n_vars, 180, seed=4)
    unit = bind(ctx, n_vars, clauses, chains, H.make_weights(seed=1))
    dev = torch.device("cuda:0")
    tdt = torch.float32 if dtype == "f32" else torch.bfloat16
    code = 0 if dtype == "f32" else 1
    rng = torch.Generator().manual_seed(0)
    for direction, rows_in, rows_out, rowptr, col, scale in (
            (0, 2 * n_vars, unit.n_clauses, unit.cl_rowptr, unit.cl_lit, unit.rev_degree_weight()),
            (1, unit.n_clauses, 2 * n_vars, unit.lit_rowptr, unit.lit_clause, unit.degree_weight())):
        x = torch.randn(chains, rows_in, feat, generator=rng).to(tdt)
        xd = x.to(dev)
        yd = torch.zeros(chains, rows_out, feat, dtype=tdt, device=dev)
        torch.cuda.synchronize()
        ctx.spmm(direction, xd.data_ptr(), yd.data_ptr(), feat, code, chains)
        ctx.synchronize()
        want = np.zeros((chains, rows_out, feat), dtype=np.float32)
        xf = x.float().numpy()
        for r in range(rows_out):
            acc = np.zeros((chains, feat), dtype=np.float32)
            for e in range(rowptr[r], rowptr[r + 1]):
                acc = acc + xf[:, col[e]]
            want[:, r] = acc * scale[r]
        got = yd.float().cpu().numpy()
        if dtype == "f32":
            np.testing.assert_array_equal(got, want)      # same summation order: bit exact
        else:
            np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-2)


def test_sampler_end_to_end(ctx, tmp_path):
    from diffusionsat_b200.sampler import DiffusionSampler
    from diffusionsat_b200.weights import save_weights
    n_vars, clauses = 5, [[-1, 2], [1, -2], [-3, 4, 5]]          # 14 models (reference utils/test_AllSolutions.py)
    cnf = tmp_path / "f.cnf"
    cnf.write_text(synth.dimacs_text(n_vars, clauses))
    wpath = tmp_path / "w.npz"
    save_weights(str(wpath), H.make_weights(seed=2))
    sampler = DiffusionSampler(str(wpath), str(cnf), context=ctx, chains_per_launch=512, seed=3,
                               max_nodes_per_batch=130)
    hist = sampler.samples(100)
    models = set(synth.enumerate_solutions(n_vars, clauses))
    assert len(models) == 14
    if sampler.last_stats["sat"] / max(sampler.last_stats["total"], 1) >= 0.005:
        assert sum(hist.values()) == 100
    assert set(hist) <= models
    # independent of launch size: same seed, different chains_per_launch -> same histogram
    sampler2 = DiffusionSampler(str(wpath), str(cnf), context=ctx, chains_per_launch=2048, seed=3,
                                max_nodes_per_batch=130)
    assert sampler2.samples(100) == hist
    # model_path as the reference's TensorFlow checkpoint directory (read without TensorFlow): same weights, same histogram
    from diffusionsat_b200.tf_checkpoint import save_querysat_checkpoint
    save_querysat_checkpoint(str(tmp_path / "tf_model"), H.make_weights(seed=2), step=5)
    sampler3 = DiffusionSampler(str(tmp_path / "tf_model"), str(cnf), context=ctx, chains_per_launch=512, seed=3,
                                max_nodes_per_batch=130)
    assert sampler3.samples(100) == hist


# ---------------------------------------------------------------- against the reference's own source
# golden vectors from tests/golden/make_model_golden.py (reference model/query_sat.py and
# satuniformity/DiffusionSampler.py executed over the torch-backed TF stand-in)
from tests.test_oracle_vs_reference_golden import DIFF_TAGS, STEP_TAGS, diff_case, step_case  # noqa: E402


@pytest.mark.parametrize("tag", STEP_TAGS)
def test_cuda_model_call_matches_reference_source(ctx, tag):
    c = step_case(tag)
    wts = H.W.init_weights(seed=c["wseed"], bias_scale=0.1)
    bind(ctx, c["n_vars"], c["clauses"], c["chains"], wts)
    pred, steps, loss = ctx.model_call(c["noise_scale"], c["noisy"], labels=c["labels"], normals=c["normals"],
                                       rounds=c["rounds"])
    assert steps[0] == c["steps_taken"]
    check("prediction vs reference source", pred, c["prediction"], 1e-3)
    assert abs(loss[0] - c["loss"]) < 1e-3 * max(1.0, abs(c["loss"]))


@pytest.mark.parametrize("tag", DIFF_TAGS)
def test_cuda_diffusion_matches_reference_source(ctx, tag):
    c = diff_case(tag)
    wts = H.W.init_weights(seed=c["wseed"], bias_scale=0.1)
    bind(ctx, c["n_vars"], c["clauses"], c["chains"], wts, group=c["chains"])
    packed, is_sat, latch, _ = ctx.sample(c["steps"], c["rounds"], uniforms=c["uniforms"], labels=c["labels"],
                                          normals=c["normals"])
    from diffusionsat_b200.sampler import unpack_assignments
    got = unpack_assignments(packed, c["n_vars"])
    want = [O.encode_assignment(c["predictions"][i * c["n_vars"]:(i + 1) * c["n_vars"]]) for i in range(c["chains"])]
    assert got == want


def test_uniformity_statistic_comparable_to_oracle(ctx):
    """Sample histogram of the CUDA path vs the oracle's on a formula with 14 models (reference
    utils/test_AllSolutions.py): same Philox noise spec on both sides, so the histograms agree up to rare
    rounding-boundary flips, and so do their chi-square statistics against the uniform distribution
    (reference utils/chi_square.py, diffusion_metrics.py:137)."""
    from diffusionsat_b200.sampler import unpack_assignments
    from diffusionsat_b200.uniformity import chi_square_vs_ideal
    n_vars, clauses = 5, [[-1, 2], [1, -2], [-3, 4, 5]]
    chains, steps, rounds, seed = 1500, 6, 3, 11
    wts = H.make_weights(seed=6)
    bind(ctx, n_vars, clauses, chains, wts, group=chains)
    packed, is_sat, _, _ = ctx.sample(steps, rounds, seed=seed)
    got = {}
    for v, s in zip(unpack_assignments(packed, n_vars), is_sat):
        if s:
            got[v] = got.get(v, 0) + 1
    # oracle with the identical noise streams (host restatement of the device Philox)
    elems = np.arange(chains * n_vars)
    uniforms = np.stack([philox.uniforms(seed, elems, t) for t in range(steps)])
    labels = np.stack([philox.labels(seed, elems, t) for t in range(steps)])
    normals = np.stack([np.stack([philox.normals(seed, elems, t, r) for r in range(rounds)]) for t in range(steps)])
    graph = O.OracleGraph.copies(n_vars, clauses, chains)
    _, final, _ = O.diffusion(steps, graph, O.weights_to_torch(wts), torch.from_numpy(uniforms),
                              torch.from_numpy(labels.astype(np.int64)), torch.from_numpy(normals), rounds)
    want = {}
    for c in range(chains):
        bits = final[c * n_vars:(c + 1) * n_vars]
        if O._satisfiable_py([bool(b) for b in bits], clauses):
            k = O.encode_assignment(bits)
            want[k] = want.get(k, 0) + 1
    models = synth.enumerate_solutions(n_vars, clauses)
    assert set(got) <= set(models) and set(want) <= set(models)
    differing = sum(abs(got.get(k, 0) - want.get(k, 0)) for k in models)
    assert differing <= 0.02 * chains                      # chains that flipped on a rounding boundary
    chi_g, _ = chi_square_vs_ideal(got, models)
    chi_o, _ = chi_square_vs_ideal(want, models)
    assert abs(chi_g - chi_o) <= 0.1 * max(chi_o, 1.0) + 5.0
