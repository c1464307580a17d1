"""The CPU oracle against golden vectors computed by the REFERENCE's own model source
(tests/golden/make_model_golden.py runs model/query_sat.py, satuniformity/DiffusionSampler.py, ... over
oracle/tf_shim.py).  This pins the oracle's control flow, tensor plumbing and op order to the reference;
what stays unpinned is the numerical behaviour of each TensorFlow kernel itself."""
import ast
import os

import numpy as np
import pytest
import torch

from diffusionsat_b200 import weights as W
from oracle import querysat_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz"), allow_pickle=False)
STEP_TAGS = ["a", "b", "c", "d"]
DIFF_TAGS = ["e", "f"]


def step_case(tag):
    g = lambda k: GOLD["step_%s_%s" % (tag, k)]
    return dict(n_vars=int(g("n_vars")), clauses=ast.literal_eval(str(g("clauses")[0])), chains=int(g("chains")),
                rounds=int(g("rounds")), noise_scale=float(g("noise_scale")), wseed=int(g("wseed")), labels=g("labels"),
                normals=g("normals"), noisy=g("noisy"), uniform=g("uniform"), prediction=g("prediction"),
                steps_taken=int(g("steps_taken")), loss=float(g("loss")))


def diff_case(tag):
    g = lambda k: GOLD["diff_%s_%s" % (tag, k)]
    return dict(n_vars=int(g("n_vars")), clauses=ast.literal_eval(str(g("clauses")[0])), chains=int(g("chains")),
                steps=int(g("steps")), rounds=int(g("rounds")), wseed=int(g("wseed")), uniforms=g("uniforms"),
                labels=g("labels"), normals=g("normals"), predictions=g("predictions"), accuracy=float(g("accuracy")))


@pytest.mark.parametrize("tag", STEP_TAGS)
def test_model_call_matches_reference_source(tag):
    c = step_case(tag)
    graph = O.OracleGraph.copies(c["n_vars"], c["clauses"], c["chains"])
    w = O.weights_to_torch(W.init_weights(seed=c["wseed"], bias_scale=0.1))
    noisy = O.randomized_rounding(torch.full((graph.n_vars, 2), 0.5), torch.from_numpy(c["uniform"]))
    np.testing.assert_array_equal(noisy.numpy(), c["noisy"])                      # randomized_rounding_tf
    out = O.model_call(graph, w, c["noise_scale"], noisy, torch.from_numpy(c["labels"].astype(np.int64)),
                       torch.from_numpy(c["normals"]), c["rounds"])
    assert out["steps_taken"] == c["steps_taken"]                                 # incl. the early exit of case c
    np.testing.assert_allclose(out["prediction"].numpy(), c["prediction"], rtol=2e-4, atol=2e-5)
    assert float(out["loss"]) == pytest.approx(c["loss"], rel=1e-4)


def test_posterior_matches_reference_source():
    x, p = torch.from_numpy(GOLD["post_x"]), torch.from_numpy(GOLD["post_p"])
    x0 = torch.stack([1 - p, p], dim=1)
    for i in range(4):
        got = O.reverse_distribution_step_theoretic(x, x0, float(GOLD["post_%d_t" % i]), 1 / 32)
        np.testing.assert_allclose(got.numpy(), GOLD["post_%d_out" % i], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", DIFF_TAGS)
def test_diffusion_loop_matches_reference_source(tag):
    c = diff_case(tag)
    graph = O.OracleGraph.copies(c["n_vars"], c["clauses"], c["chains"])
    w = O.weights_to_torch(W.init_weights(seed=c["wseed"], bias_scale=0.1))
    acc, final, _ = O.diffusion(c["steps"], graph, w, torch.from_numpy(c["uniforms"]),
                                torch.from_numpy(c["labels"].astype(np.int64)), torch.from_numpy(c["normals"]), c["rounds"])
    np.testing.assert_array_equal(final, c["predictions"])
    assert acc == pytest.approx(c["accuracy"])


UNION = np.load(os.path.join(os.path.dirname(__file__), "golden", "union_golden.npz"), allow_pickle=False)


@pytest.mark.parametrize("tag", ["u", "v"])
def test_union_batch_model_call_matches_reference_source(tag):
    """Disjoint unions of DIFFERENT formulas (tests/golden/make_union_golden.py): graphs of unequal size in one batch, as in
    training batches and predict_step -- per-graph PairNorm statistics, per-graph logit-map choice, whole-batch early exit."""
    g = lambda k: UNION["%s_%s" % (tag, k)]
    formulas = ast.literal_eval(str(g("formulas")[0]))
    graph = O.OracleGraph.from_formulas(formulas)
    assert graph.n_vars == sum(n for n, _ in formulas)
    w = O.weights_to_torch(W.init_weights(seed=int(g("wseed")), bias_scale=0.1))
    noisy = O.randomized_rounding(torch.full((graph.n_vars, 2), 0.5), torch.from_numpy(g("uniform")))
    np.testing.assert_array_equal(noisy.numpy(), g("noisy"))
    out = O.model_call(graph, w, float(g("noise_scale")), noisy, torch.from_numpy(g("labels").astype(np.int64)),
                       torch.from_numpy(g("normals")), int(g("rounds")))
    assert out["steps_taken"] == int(g("steps_taken"))
    np.testing.assert_allclose(out["prediction"].numpy(), g("prediction"), rtol=2e-4, atol=2e-5)
    assert float(out["loss"]) == pytest.approx(float(g("loss")), rel=1e-4)


def test_predict_step_tries_bookkeeping_matches_reference_source():
    """``QuerySAT.predict_step`` with ``prediction_tries = 3`` (reference model/query_sat.py:424-451): the host bookkeeping
    around the model calls -- per-try ``is_graph_sat``, "newly solved" clipping, per-variable masks, zeros for graphs never
    solved -- against the reference's own method run with the same three stubbed logit vectors."""
    from diffusionsat_b200 import graph as G
    from diffusionsat_b200.query_sat import QuerySAT
    formulas = ast.literal_eval(str(UNION["p_formulas"][0]))
    union = G.build_union_graph(formulas)
    coo, shape = union.reference_coo(1)
    vg = np.repeat(np.arange(len(formulas)), [n for n, _ in formulas])
    cg = np.repeat(np.arange(len(formulas)), [len(c) for _, c in formulas])
    model = QuerySAT.__new__(QuerySAT)                      # no GPU context: only the bookkeeping is under test
    model.prediction_tries = int(UNION["p_logits"].shape[0])
    feed = iter(UNION["p_logits"])
    model.call = lambda *a, **k: (next(feed).copy(), np.float32(0.25), 7)
    res = model.predict_step((coo, shape), cg, vg)
    np.testing.assert_array_equal(res["prediction"], UNION["p_prediction"])
    assert res["steps_taken"] == int(UNION["p_steps_taken"]) and float(res["loss"]) == float(UNION["p_loss"])
    never = slice(union.n_vars - 2, union.n_vars)           # the unsatisfiable last formula keeps zeros
    assert not res["prediction"][never].any() and res["prediction"].any()
