"""N-GPU run (torchrun, NCCL) of the scenario callers: every rank must report the numbers of the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/multi_gpu_callers.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib  # noqa: E402
from monte_carlo_option_simulator_b200.dist import TorchComm  # noqa: E402
from monte_carlo_option_simulator_b200.risk import HedgingBacktest, StressTestEngine  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
comm = TorchComm()
h = _lib.Handle(local)
p = SVJParams()


def both(f):
    """(sharded result, seconds, single-GPU result computed on this rank)"""
    f(comm)
    dist.barrier()
    t0 = time.perf_counter()
    r = f(comm)
    dt = time.perf_counter() - t0
    return r, dt, f(None)


def close(a, b, tol=1e-9):
    if isinstance(a, dict):
        return all(close(a[k], b[k], tol) for k in a)
    if isinstance(a, (list, tuple)):
        return all(close(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, float) and a != a:
        return b != b
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


res = {}
res["stress"] = both(lambda c: StressTestEngine(p, num_paths=2_000_000, seed=42, handle=h, comm=c).full_stress_report(22500.0, 22500.0, 0.25))
res["hedge"] = both(lambda c: HedgingBacktest(p, seed=42, handle=h, comm=c).run_backtest(22500.0, 22500.0, 0.25))
res["price_many"] = both(lambda c: MonteCarloEngine(p, 4_000_000, 252, 7, handle=h, comm=c).price_many(
    [22500.0, 21000.0, 24000.0], 22500.0, [0.25, 0.5, 1.0], [True, False, True]))
res["qmc"] = both(lambda c: MonteCarloEngine(SVJParams.gbm(0.3), 1 << 20, 250, 42, rng="sobol", handle=h, comm=c).price(2500.0, 2500.0, 1.0))
for name, (r, dt, single) in res.items():
    ok = close(r, single, 1e-7 if name != "hedge" else 1e-9)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name:10s} world {comm.world}: {dt * 1e3:8.2f} ms, equals the single-GPU result on every rank: {bool(flag.item())}", flush=True)
    assert flag.item() == 1.0, name
dist.barrier()
dist.destroy_process_group()
h.close()
if rank == 0:
    print("MULTI-GPU CALLERS OK")
