"""Wall-clock of the reference's own small case (BASELINE configs[0]: 3-SAT n=30, samples(256)) through the public API."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionsat_b200 import synth, weights
from diffusionsat_b200.sampler import DiffusionSampler
n, clauses = synth.random_3sat(30, seed=0)
d = tempfile.mkdtemp()
cnf = os.path.join(d, "f.cnf"); open(cnf, "w").write(synth.dimacs_text(n, clauses))
wp = os.path.join(d, "w.npz"); weights.save_weights(wp, weights.init_weights(seed=1234))
for prec in ("bf16", "fp32"):
    t0 = time.perf_counter(); s = DiffusionSampler(wp, cnf, precision=prec, seed=1); t1 = time.perf_counter()
    for rep in range(3):
        t2 = time.perf_counter(); hist = s.samples(256); t3 = time.perf_counter()
        print("%s: ctor %.3f s, samples(256) call %d: %.3f s, sat %d of %d chains, distinct %d" %
              (prec, t1 - t0, rep, t3 - t2, s.last_stats["sat"], s.last_stats["total"], len(hist)))
