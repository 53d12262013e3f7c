"""Given-normals kernel on DEVICE-resident normals (torch tensors): GB/s for SVJ / Heston / GBM parameter sets."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from monte_carlo_option_simulator_b200 import SVJParams, _lib
h = _lib.Handle(0)
for n, steps in ((1_000_000, 250), (4_000_000, 64)):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    d = [torch.randn(n, steps, dtype=torch.float64, device="cuda", generator=g) for _ in range(2)]
    d.append(torch.rand(n, steps, dtype=torch.float64, device="cuda", generator=g)); d.append(torch.randn(n, steps, dtype=torch.float64, device="cuda", generator=g))
    S = torch.empty(n, dtype=torch.float64, device="cuda"); V = torch.empty_like(S)
    for name, p, narr in (("svj", SVJParams(), 4), ("heston", SVJParams(lambda_j=0.0), 2), ("gbm", SVJParams.gbm(0.3), 1)):
        sp = _lib.to_params(p)
        out = []
        ref = None
        for ilp in ("0", "1"):                       # B200MC_GN_ILP: three-sweep tile walk (A/B)
            for tile in ("8", "16"):
                os.environ["B200MC_GN_ILP"], os.environ["B200MC_GN_TILE"] = ilp, tile
                best = 1e9
                for _ in range(4):
                    h.timer_begin()
                    h._check(h.lib.b200mc_simulate_given_normals_dev(h.h, C.byref(sp), 2500.0, 1.0, n, steps, *(C.c_void_p(t.data_ptr()) for t in d), 0,
                                                                    C.c_void_p(S.data_ptr()), C.c_void_p(V.data_ptr()), None))
                    best = min(best, h.timer_end())
                out.append(f"{'sweeps' if ilp == '1' else 'steps '} TS={tile}: {best:.3f} ms {narr * n * steps * 8 / best / 1e6:.0f} GB/s")
                if ref is None:
                    ref = (S.clone(), V.clone())
                else:
                    assert torch.equal(S, ref[0]) and torch.equal(V, ref[1]), "variants disagree"
        os.environ.pop("B200MC_GN_ILP"); os.environ.pop("B200MC_GN_TILE")
        print(f"n={n} steps={steps} {name:7s}: " + " | ".join(out) + "  (needed input bytes; all variants bitwise equal)", flush=True)
    del d
