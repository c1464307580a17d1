"""Host-side logic of the sampler that needs no GPU: the reference's stop rules applied to a launch's SAT flags
(satuniformity/DiffusionSampler.py:243-307), histogram table merging, key -> int conversion, checkpoint lookup."""
import numpy as np
import pytest

from diffusionsat_b200 import dist as D
from diffusionsat_b200.sampler import DiffusionSampler, consume_batches


def _reference_loop(is_sat, batch, need, total, sat_total, rate=0.005):
    """The per-sample loop of the reference, one reference batch at a time."""
    used = sat = 0
    for b0 in range(0, len(is_sat), batch):
        if need == 0:
            break
        if total > 0 and sat_total / total < rate:          # :261-263
            return used, sat, True
        for i in range(b0, min(b0 + batch, len(is_sat))):
            total += 1
            used = i + 1
            if is_sat[i]:                                   # :297-303
                sat_total += 1
                sat += 1
                need -= 1
                if need == 0:                               # :305-307
                    break
    return used, sat, False


def test_consume_batches_equals_the_reference_loop():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        n, b = int(rng.integers(1, 60)), int(rng.integers(1, 12))
        flags = (rng.random(n) < rng.choice([0.0, 0.003, 0.3, 0.9])).astype(np.uint8)
        need, tot = int(rng.integers(1, 40)), int(rng.integers(0, 300))
        st = int(rng.integers(0, tot + 1)) if tot and rng.random() < 0.5 else 0
        assert consume_batches(flags, b, need, tot, st) == _reference_loop(flags, b, need, tot, st)


def test_merge_tables_and_key_conversion():
    rng = np.random.default_rng(1)
    words, n_bits = 2, 100
    pool = rng.integers(0, 2**63, size=(9, words), dtype=np.uint64)
    pool[:, 1] &= np.uint64((1 << (n_bits - 64)) - 1)
    want = {}
    tables = []
    for _ in range(4):
        pick = rng.integers(0, len(pool), 30)
        packed = pool[pick]
        sat = (rng.random(30) < 0.7).astype(np.uint8)
        tables.append(D.local_histogram(packed, sat))
        for row, s in zip(packed, sat):
            if s:
                k = int(row[0]) | (int(row[1]) << 64)
                want[k] = want.get(k, 0) + 1
    keys, counts = D.merge_tables(tables, words)
    ints = D.keys_to_ints(keys, n_bits)
    assert ints == sorted(ints)                             # ascending by the encoded integer
    assert dict(zip(ints, counts.tolist())) == want
    assert D.table_to_dict(keys, counts, n_bits) == want
    k0, c0 = D.merge_tables([], words)
    assert k0.shape == (0, words) and c0.shape == (0,)


def test_missing_checkpoint_is_random_init_but_a_broken_one_raises(tmp_path, capsys):
    s = DiffusionSampler.__new__(DiffusionSampler)
    w = s._prepare_checkpoints(str(tmp_path / "nothing_here"))
    assert "Checkpoint not found!" in capsys.readouterr().out
    assert w.feature_maps == 128
    bad = tmp_path / "weights.npz"
    bad.write_bytes(b"not a zip archive")
    with pytest.raises(Exception):
        s._prepare_checkpoints(str(bad))


def test_launch_sizing_rules():
    """Launches hold whole reference batches; the first one assumes a 50 % SAT rate and at least four batches, later ones
    follow the observed rate; a remainder smaller than a batch rides along; an explicit launch size is rounded down."""
    class FakeCtx:
        graph, chains = None, 0
    s = DiffusionSampler.__new__(DiffusionSampler)
    s.ctx, s.unit = FakeCtx(), object()
    s.batch_chains, s.clauses, s.n_vars, s.chains_per_launch = 103, [[1]] * 133, 30, None
    assert s._launch_chains(256) == 6 * 103                      # 256 / 0.5 * 1.15 -> 6 batches
    assert s._launch_chains(10) == 4 * 103                       # never fewer than four batches
    assert s._launch_chains(256, sat_rate=0.9) == 4 * 103
    assert s._launch_chains(10 ** 9) % 103 == 0 and s._launch_chains(10 ** 9) * 133 <= 900_000     # memory cap
    s.ctx.graph, s.ctx.chains = s.unit, 7 * 103
    assert s._launch_chains(256) == 7 * 103                      # same shape as the previous launch is kept
    s.ctx.graph = None
    s.chains_per_launch = 4096
    s.batch_chains = 31
    assert s._launch_chains(10 ** 9) == 4092
    assert s._launch_chains(10 ** 9, chains_left=4096) == 4096   # 4 chains left over ride along as a partial group
    assert s._launch_chains(10 ** 9, chains_left=10000) == 4092
    assert s._launch_chains(10 ** 9, chains_left=20) == 20


def test_stop_rules_and_histogram_match_the_reference_samples_loop():
    """The sampler's host logic -- reference batches consumed in order with ``consume_batches``, histogram of the consumed
    satisfying chains -- replayed on the assignments that the REFERENCE's own ``samples()`` loop was fed
    (tests/golden/make_samples_golden.py: satuniformity/DiffusionSampler.py:229-311 over the TF stand-in, ``diffusion()``
    stubbed).  The histogram must not depend on how many reference batches a launch holds."""
    import ast
    import os
    from diffusionsat_b200.variable_assignment import VariableAssignment
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "samples_golden.npz"))
    n_vars, graphs = int(gold["n_vars"]), int(gold["graphs"])
    clauses = ast.literal_eval(str(gold["clauses"][0]))

    def encode(bits):
        asgn = VariableAssignment(clauses=clauses)
        asgn.assign_all_from_bit_list([float(b) for b in bits])
        return int(asgn), asgn.satisfiable()

    for name in ("mixed", "unsat_first"):
        batches = gold[name + "_batches"]
        chains = batches.reshape(len(batches) * graphs, n_vars)
        coded = [encode(row) for row in chains]
        for n_samples in (1, 5, 17, 1000):
            want = dict(zip(gold["%s_%d_keys" % (name, n_samples)].tolist(), gold["%s_%d_counts" % (name, n_samples)].tolist()))
            for per_launch in (1, 2, 5, len(batches)):
                hist, total, sat_total, need, pos, stop = {}, 0, 0, n_samples, 0, False
                while need > 0 and not stop and pos < len(coded):
                    launch = coded[pos:pos + per_launch * graphs]
                    flags = np.array([s for _, s in launch], dtype=np.uint8)
                    used, sat_used, stop = consume_batches(flags, graphs, need, total, sat_total)
                    for key, sat in launch[:used]:
                        if sat:
                            hist[key] = hist.get(key, 0) + 1
                    total, sat_total, need, pos = total + used, sat_total + sat_used, need - sat_used, pos + per_launch * graphs
                assert hist == want, (name, n_samples, per_launch)
                assert stop == bool(gold["%s_%d_aborted" % (name, n_samples)]), (name, n_samples, per_launch)
    assert sum(gold["mixed_5_counts"]) == 5 and not len(gold["unsat_first_1000_keys"])
