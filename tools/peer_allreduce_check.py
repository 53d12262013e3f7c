"""torchrun check of b200mc_peer_allreduce (one-shot all-reduce over NVLink peer memory) against NCCL: same sums on every
rank for many epochs and sizes, bitwise equal across ranks, and the device time of both for the 17-double message.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        tools/peer_allreduce_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import _lib  # noqa: E402
from monte_carlo_option_simulator_b200.dist import PeerComm  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
world = dist.get_world_size()
h = _lib.Handle(local)
h.set_stream(torch.cuda.current_stream().cuda_stream)
comm = PeerComm(h)

g = np.random.default_rng(100 + rank)
ok = True
for it in range(300):
    n = int(np.random.default_rng(it).choice([1, 17, 136, 357, 1088, 4352]))
    x = torch.from_numpy(g.standard_normal(n) * 10.0 ** g.integers(-3, 6)).cuda()
    ref = x.clone()
    dist.all_reduce(ref)
    h.peer_allreduce(x.data_ptr(), n)
    torch.cuda.synchronize()
    ok &= bool(torch.allclose(x, ref, rtol=1e-13, atol=1e-9))
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    ok &= all(bool(torch.equal(gathered[0], t)) for t in gathered)          # bitwise identical on every rank
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
assert flag.item() == 1.0

x = torch.ones(17, dtype=torch.float64, device="cuda")
res = {}
for name, f in (("nccl", lambda: dist.all_reduce(x)), ("peer", lambda: h.peer_allreduce(x.data_ptr(), 17))):
    for _ in range(20):
        f()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        x.fill_(1.0)
        f()
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 200 * 1e3
if rank == 0:
    print(f"world {world}: 300 epochs of random sizes equal NCCL's sums and are bitwise identical across ranks: {ok}")
    print(f"17 doubles, back to back on one stream (incl. a fill kernel): NCCL {res['nccl']:.1f} us, peer memory {res['peer']:.1f} us")
    print("PEER ALLREDUCE OK")
dist.barrier()
h.close()
dist.destroy_process_group()
