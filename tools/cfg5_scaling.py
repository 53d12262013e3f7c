"""BASELINE cfg5 (strong scaling): ONE European call priced with n_total in {1e8, 1e9, 1e10} paths x 250 steps, seed 42,
path range [g n/G, (g+1) n/G) on GPU g, through MonteCarloEngine.price with the peer-memory exchange.  Prints seconds,
path-steps/s, the price with its standard error and the relative deviation from Black-Scholes.

    python tools/cfg5_scaling.py                                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/cfg5_scaling.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib, bs_price  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
comm = None
h = _lib.Handle(local)
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    from monte_carlo_option_simulator_b200.dist import PeerComm
    comm = PeerComm(h)
p = SVJParams.gbm(0.3, r=0.065)
bs = bs_price(2500.0, 2500.0, 1.0, p.r, p.q, 0.3, True)
for n_total in (10 ** 8, 10 ** 9, 10 ** 10):
    e = MonteCarloEngine(p, n_total, 250, 42, use_sobol=False, use_antithetic=False, use_control_variate=False, rng="philox",
                         handle=h, comm=comm)
    e.price(2500.0, 2500.0, 1.0)                     # warm-up at full size
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = e.price(2500.0, 2500.0, 1.0)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"G={world} n_total={n_total:.0e}: {dt * 1e3:9.2f} ms  {n_total * 250 / dt:.3e} path-steps/s   price {r['price']:.5f} +- "
              f"{r['std_error']:.5f}  BS {bs:.5f}  rel {abs(r['price'] - bs) / bs:.1e}  z {(r['price'] - bs) / r['std_error']:+.2f}   "
              f"spot-CV {r['price_cv_spot']:.5f} +- {r['std_error_cv_spot']:.5f}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
h.close()
