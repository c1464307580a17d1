"""Generate tests/golden/helper_golden.npz: the module-level helpers of the REFERENCE's model/query_sat.py
(``randomized_rounding_tf`` :55-60, ``distribution_at_time`` :66-68, ``add_t_emb`` :70-74, ``construct_training_input`` :76-82)
executed over oracle/tf_shim.py with the uniform draws injected.  diffusionsat_b200/query_sat.py restates them in numpy
(they build the model input outside ``diffusion_step``: ``call`` without a noisy input, ``predict_step``).

Run in the build container only:  python tests/golden/make_helper_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_model_golden as M  # noqa: E402  (installs the TF stand-in and imports the reference's modules)
import model.query_sat as ref_qs  # noqa: E402  (the reference's module)

from oracle import tf_shim  # noqa: E402

OUT = os.path.join(HERE, "helper_golden.npz")


def main():
    rng = np.random.default_rng(31337)
    n = 64
    out = {}
    x = rng.random((n, 2)).astype(np.float32)
    x[:4, 0] = [0.0, 1.0, 0.5, 0.25]
    u = rng.random((n, 1)).astype(np.float32)
    u[:4, 0] = [0.0, 0.0, 0.5, 0.75]                        # floor(x0 + u) on the boundaries
    tf_shim.NOISE.clear()
    tf_shim.NOISE.uniforms.append(torch.from_numpy(u))
    out["rr_x"], out["rr_u"] = x, u
    out["rr_out"] = ref_qs.randomized_rounding_tf(torch.from_numpy(x)).numpy()
    for i, t in enumerate([0.0, 0.3, 1.0]):
        out["dat_%d_t" % i] = np.float32(t)
        out["dat_%d_out" % i] = ref_qs.distribution_at_time(torch.from_numpy(x), t).numpy()
    out["emb_out"] = ref_qs.add_t_emb(torch.from_numpy(x), 0.625).numpy()
    bits = rng.integers(0, 2, n).astype(np.int32)
    for i, t in enumerate([1.0, 0.5, 1 / 32]):
        tf_shim.NOISE.clear()
        tf_shim.NOISE.uniforms.append(torch.from_numpy(u))
        out["cti_%d_t" % i] = np.float32(t)
        out["cti_%d_out" % i] = ref_qs.construct_training_input(torch.from_numpy(bits.astype(np.int64)), torch.tensor(t)).numpy()
    out["cti_bits"] = bits
    out["t_power"] = np.float32(ref_qs.t_power)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; t_power", ref_qs.t_power)


if __name__ == "__main__":
    main()
