"""Golden vectors that close the north star's deterministic-mode loop with the REFERENCE itself (not the oracle):

  stage 1 (GPU box):        python tests/golden/make_fused_golden.py --dump gpurun_out/fused_draws.npz
      dumps the exact Philox/Box-Muller draws the fused kernels consume for a few (mode, seed, offset, shape) cases;
  stage 2 (dev container):  python tests/golden/make_fused_golden.py --reference gpurun_out/fused_draws.npz
      feeds those draws to the reference's own _simulate_svj_paths_numba (imported from /root/reference) and writes
      tests/golden/fused_golden.npz = draws + the reference's S_T, v_T, antithetic S_T, path matrix.

tests/test_gpu_parity.py::test_fused_modes_against_the_reference_itself then checks, on a GPU, that (a) the library
still produces bit-identical draws and (b) the fused kernels reproduce the reference's outputs (<= 1e-6 fp64, 1e-4 fp32).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    # name: (params, S0, T, steps, n, seed, path_offset)
    "gbm": (dict(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0),
            2500.0, 1.0, 250, 96, 42, 0),
    "detvar": (dict(kappa=3.0, theta=0.05, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0),
               2500.0, 0.5, 125, 96, 7, 1_000_003),
    "heston": (dict(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.012),
               22500.0, 0.25, 63, 128, 11, 2 ** 32 - 100),
    "svj": (dict(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.10, r=0.065, q=0.012),
            22500.0, 0.25, 63, 128, 42, 12345),
    "jumpy": (dict(kappa=2.0, theta=0.04, xi=0.3, rho=-0.5, v0=0.05, lambda_j=12.0, mu_j=-0.08, sigma_j=0.2, r=0.03, q=0.0),
              100.0, 1.0, 100, 128, 3, 0),
}


def dump(path):
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    h = _lib.Handle(0)
    out = {}
    for name, (pd, S0, T, steps, n, seed, off) in CASES.items():
        p = SVJParams(**pd)
        stream = _lib.select_stream(p, T, steps)
        for i, w in enumerate(("Z1", "Z2", "Zj", "Zjs")):
            out[f"{name}_{w}"] = h.dump_normals(seed, n, steps, stream, i, path_offset=off, jump_prob=p.lambda_j * (T / steps))
        out[f"{name}_stream"] = np.array([stream])
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("Z1")})


def reference(path):
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.path.insert(0, "/root/reference")
    from engine.monte_carlo import _simulate_svj_paths_numba as ref_kernel          # the reference itself
    d = dict(np.load(path))
    out = dict(d)
    for name, (pd, S0, T, steps, n, seed, off) in CASES.items():
        Z1, Z2, Zj, Zjs = (np.ascontiguousarray(d[f"{name}_{w}"]) for w in ("Z1", "Z2", "Zj", "Zjs"))
        args = (pd["r"], pd["q"], T, pd["kappa"], pd["theta"], pd["xi"], pd["rho"], pd["lambda_j"], pd["mu_j"], pd["sigma_j"])
        S, v, paths = ref_kernel(float(S0), pd["v0"], *args, Z1, Z2, Zj, Zjs, steps, True)
        Sa, _, _ = ref_kernel(float(S0), pd["v0"], *args, -Z1, -Z2, Zj, -Zjs, steps, False)
        Su, _, _ = ref_kernel(float(S0), pd["v0"] + 0.01, *args, Z1, Z2, Zj, Zjs, steps, False)
        out[f"{name}_ref_S"], out[f"{name}_ref_v"], out[f"{name}_ref_paths"] = S, v, paths
        out[f"{name}_ref_S_anti"], out[f"{name}_ref_S_v0up"] = Sa, Su
    dst = os.path.join(HERE, "fused_golden.npz")
    np.savez_compressed(dst, **out)
    with open(os.path.join(HERE, "fused_golden_cases.json"), "w") as f:
        json.dump({k: dict(params=v[0], S0=v[1], T=v[2], steps=v[3], n=v[4], seed=v[5], path_offset=v[6]) for k, v in CASES.items()}, f, indent=1)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    {"--dump": dump, "--reference": reference}[sys.argv[1]](sys.argv[2])
