"""BASELINE cfg4 on N GPUs (torchrun): 4M paths x 251 fp32 stored SHARDED (every rank keeps rows [lo, hi) of the global
path index in its own HBM, no exchange), discounted option P&L formed on the device from the last column, VaR / CVaR /
tail metrics over the sharded P&L vector (distributed radix select: only histograms and a few sums cross the ranks).
Every rank must report the metrics of the single-GPU pipeline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/cfg4_multi_gpu.py
"""
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402
from monte_carlo_option_simulator_b200.dist import PeerComm, TorchComm, sharded_generate_paths  # noqa: E402
from monte_carlo_option_simulator_b200.risk import KEYS, compute_risk_metrics_sharded  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
world = dist.get_world_size()
h = _lib.Handle(local)
h.set_stream(torch.cuda.current_stream().cuda_stream)
p = SVJParams.gbm(0.3, r=0.065)
N, STEPS, K, S0, T = 4_000_000, 250, 2500.0, 2500.0, 1.0
disc = math.exp(-p.r * T)

# the single-GPU pipeline on this rank: all 4M rows, P&L from the last column, radix select on one device
full = torch.empty(N * (STEPS + 1), dtype=torch.float32, device="cuda")
pnl1 = torch.empty(N, dtype=torch.float64, device="cuda")
h.generate_paths(p, S0, T, STEPS, N, 42, 0, np.float32, out_dev=full.data_ptr())
h.option_pnl(full.data_ptr() + STEPS * 4, N, K, True, disc, 374.0712289657911, pnl1.data_ptr(), dtype_in=np.float32, stride=STEPS + 1)
single = dict(zip(KEYS, h.risk_metrics(pnl1.data_ptr(), 0.99, n=N, dtype=np.float64)))
del full, pnl1
class HostLoop(TorchComm):
    """NCCL + the host-driven select (risk_begin / _hist / _finish with NumPy all-reduces between the passes)"""


for name, comm in (("nccl, host-driven select", HostLoop()), ("peer memory, device-side select", PeerComm(h))):
    lo, hi = 0, 0
    n_loc = N // world + 1
    mat = torch.empty(n_loc * (STEPS + 1), dtype=torch.float32, device="cuda")      # allocated outside the timed region
    pnl = torch.empty(n_loc, dtype=torch.float64, device="cuda")
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lo, hi, _ = sharded_generate_paths(h, comm, p, S0, T, STEPS, N, 42, 0, np.float32, out_dev=mat.data_ptr())
        h.option_pnl(mat.data_ptr() + STEPS * 4, hi - lo, K, True, disc, 374.0712289657911, pnl.data_ptr(), dtype_in=np.float32,
                     stride=STEPS + 1)
        torch.cuda.synchronize()          # the launches above are asynchronous: without this t1 is enqueue time
        t1 = time.perf_counter()
        got = compute_risk_metrics_sharded((pnl.data_ptr(), hi - lo, np.float64), 0.99, comm=comm, handle=h)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    ok = all((math.isnan(single[k]) and math.isnan(got[k])) or abs(got[k] - single[k]) <= 1e-9 * max(1.0, abs(single[k]))
             for k in got)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world} exchange {name}: paths + P&L {1e3 * (t1 - t0):.2f} ms (rows [{lo}, {hi}) on rank 0), sharded tail metrics "
              f"{1e3 * (t2 - t1):.2f} ms; VaR {got['var']:.6f} CVaR {got['cvar']:.6f} tail index {got['tail_index']:.4f}; equals the "
              f"single-GPU pipeline on every rank: {bool(flag.item())}", flush=True)
    assert flag.item() == 1.0
    # a heavy-tailed vector in uneven shards (rank 0 holds a third, the last rank possibly nothing of the remainder's tail)
    x = np.random.default_rng(3).standard_t(3, size=3_000_001) * 0.01
    cuts = [0, x.size // 3] + [x.size // 3 + (x.size - x.size // 3) * (r + 1) // max(world - 1, 1) for r in range(world - 1)]
    cuts = (cuts + [x.size])[:world + 1]
    cuts[-1] = x.size
    want = dict(zip(KEYS, h.risk_metrics(x, 0.99)))
    got = compute_risk_metrics_sharded(x[cuts[rank]:cuts[rank + 1]], 0.99, comm=comm, handle=h)
    ok = all(abs(got[k] - want[k]) <= 1e-10 * max(1.0, abs(want[k])) for k in got)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"   Student-t(3) x 3,000,001 in shards {np.diff(cuts).tolist()}: VaR {got['var']:.6f} CVaR {got['cvar']:.6f} tail index "
              f"{got['tail_index']:.4f} kurtosis {got['kurtosis']:.2f}; equals the single-GPU metrics on every rank: {bool(flag.item())}", flush=True)
    assert flag.item() == 1.0
    # an EMPTY shard on the last rank (a rank without data still takes part in every exchange)
    y = x[:100_001]
    mine = y if rank == 0 else (y[:0] if rank == world - 1 else y[:0])
    want = dict(zip(KEYS, h.risk_metrics(y, 0.95)))
    got = compute_risk_metrics_sharded(mine, 0.95, comm=comm, handle=h)
    ok = all(abs(got[k] - want[k]) <= 1e-10 * max(1.0, abs(want[k])) for k in got)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"   all data on rank 0, empty shards elsewhere: equals the single-GPU metrics on every rank: {bool(flag.item())}", flush=True)
    assert flag.item() == 1.0
if rank == 0:
    print("CFG4 MULTI-GPU OK")
dist.barrier()
dist.destroy_process_group()
h.close()
