import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from monte_carlo_option_simulator_b200 import SVJParams, _lib
from oracle import oracle as O
h = _lib.Handle(0)
cases = {
 "gbm sigma=0.3 T=1 250": (SVJParams.gbm(0.3), 1.0, 250),
 "gbm sigma=1.0 T=5 1260": (SVJParams.gbm(1.0), 5.0, 1260),
 "heston default T=1 252": (SVJParams(lambda_j=0.0), 1.0, 252),
 "heston xi=2 T=5 1260": (SVJParams(xi=2.0, lambda_j=0.0, kappa=1.0), 5.0, 1260),
 "svj default T=1 252": (SVJParams(), 1.0, 252),
 "svj lambda=20 sigma_j=0.3 T=2 504": (SVJParams(lambda_j=20.0, sigma_j=0.3), 2.0, 504),
}
for name, (p, T, steps) in cases.items():
    n = 20000
    stream = _lib.select_stream(p, T, steps)
    Z = [h.dump_normals(11, n, steps, stream, w, jump_prob=p.lambda_j * (T / steps)) for w in range(4)]
    want, vwant, _ = O._sim(O.as_params(p), 100.0, T, *Z, steps)
    S32, A32, V32 = h.simulate_terminal(p, 100.0, T, steps, n, 11, _lib.ANTITHETIC, np.float32, 0, True, True)
    wantA = O._sim(O.as_params(p), 100.0, T, -Z[0], -Z[1], Z[2], -Z[3], steps)[0]
    paths32 = h.generate_paths(p, 100.0, T, steps, 2000, 11, 0, np.float32)
    pw = O._sim(O.as_params(p), 100.0, T, *[z[:2000] for z in Z], steps, record=True)[2]
    e1 = np.max(np.abs(S32 / want - 1)); e2 = np.max(np.abs(A32 / wantA - 1)); e3 = np.max(np.abs(paths32 / pw - 1))
    S64 = h.simulate_terminal(p, 100.0, T, steps, n, 11, _lib.FP64, np.float64)[0]
    p64 = h.generate_paths(p, 100.0, T, steps, 2000, 11, _lib.FP64, np.float64)
    print(f"{name:36s} fp64 state: S_T {np.max(np.abs(S64 / want - 1)):.1e} path matrix {np.max(np.abs(p64 / pw - 1)):.1e}")
    print(f"{name:36s} max rel err fp32: S_T {e1:.2e}  twin {e2:.2e}  path matrix {e3:.2e}   (min S_T {want.min():.3g}, max {want.max():.3g})")
