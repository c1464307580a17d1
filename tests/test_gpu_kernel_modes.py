"""The whole-MLP kernels have several instantiations picked by environment switches that libdsat reads once per process
("Experiment switches").  The default plan is what every other GPU test exercises; this file re-runs the
fused-vs-per-layer-vs-fp32 parity test of tests/test_gpu_tcgen05.py in fresh processes under the other plans, so that a
switch that is off by default (CTA pair) or a fallback (split mode off, whole bias array in shared memory) cannot rot."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODES = {
    "split_off": {"DSAT_SPLIT_MODE": "0"},                      # literal MLP through fused_mlp_kernel (one tile, 512 TMEM columns)
    "split_full_bias": {"DSAT_SPLIT_MODE": "3", "DSAT_PAIR_MODE": "4"},   # split mode, single CTA, two weight slots, biases resident
    "cta_pair": {"DSAT_PAIR_MODE": "31", "DSAT_SPLIT_MODE": "0"},   # cta_group::2 instantiation for all five MLPs
    "no_pair": {"DSAT_PAIR_MODE": "0"},                         # literal, clause and update MLPs through the single-CTA kernels too
    "one_tile_at_a_time": {"DSAT_PING_PONG": "0", "DSAT_A_RING": "0", "DSAT_PAIR_MODE": "0"},   # resident input, no ping-pong
}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", sorted(MODES))
def test_fused_mlp_parity_under_other_plans(mode):
    env = dict(os.environ)
    env.update(MODES[mode])
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_tcgen05.py"), "-x", "-q", "-m", "gpu",
           "-k", "fused_mlp_kernels_match_per_layer_kernels", "-p", "no:cacheprovider"]
    proc = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, "mode %s (%s):\n%s" % (mode, MODES[mode], proc.stdout[-3000:])
    assert "2 passed" in proc.stdout
