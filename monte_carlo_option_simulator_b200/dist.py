"""Multi-GPU sharding of the fused pricer: one process per GPU, paths split into contiguous ranges of the GLOBAL
path index (Philox counters are derived from that index, so 1/2/4/8-GPU runs draw identical numbers), and ONE
all-reduce of the per-strike sum vectors (17 doubles per strike).  Nothing in the reference corresponds to this
(it is single-process); SURVEY.md section 8(e)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ._lib import NSUMS


def shard_range(n_paths: int, rank: int, world: int):
    """[lo, hi) of the global path index owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n_paths), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Comm:
    """Minimal communicator interface: rank, world, allreduce_sum(np.float64 array) -> np.float64 array."""
    rank = 0
    world = 1

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        return a


class TorchComm(Comm):
    """torch.distributed communicator.  With the NCCL backend the sums are reduced on the device."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        if self.backend == "nccl":
            t = t.cuda()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_device(self, handle, nelem: int, fill) -> np.ndarray:
        """NCCL only: `fill(ptr)` enqueues, on the handle's stream, the kernels that write `nelem` doubles at the device
        pointer `ptr`; the buffer is all-reduced where it is and read back once (no host round trip between the kernel
        and the collective).  fill=None leaves the buffer zero (a rank without work)."""
        import torch
        dev = torch.device("cuda", handle.device)
        buf = getattr(self, "_buf", None)
        if buf is None or buf.numel() < nelem or buf.device != dev:
            buf = self._buf = torch.zeros(max(nelem, 4096), dtype=torch.float64, device=dev)
        view = buf[:nelem]
        cur = torch.cuda.current_stream(dev)
        if fill is None:
            view.zero_()
        else:
            if handle.stream != cur.cuda_stream:
                cur.synchronize()                   # the buffer may still feed the previous call's read-back
            fill(view.data_ptr())
            if handle.stream != cur.cuda_stream:
                handle.synchronize()                # order the collective after the kernels
        self._dist.all_reduce(view, op=self._dist.ReduceOp.SUM, group=self.group)
        return view.cpu().numpy()


class PeerComm(TorchComm):
    """TorchComm whose exchange of the per-strike sums is the one-shot all-reduce over NVLink peer memory of
    csrc/peer.cu (b200mc_peer_allreduce) instead of an NCCL call.  torch.distributed is used once, to all-gather the CUDA
    IPC handles of the exchange buffers; vectors that do not fit the exchange buffer, and host arrays, go through
    TorchComm's paths.  One PeerComm per (handle, process group)."""

    def __init__(self, handle, group=None):
        super().__init__(group)
        self.handle = handle
        # Every step is agreed on by all ranks before the next collective, so that a rank on which CUDA IPC is not
        # available (e.g. peers hidden by CUDA_VISIBLE_DEVICES, no P2P) makes ALL ranks raise instead of leaving the
        # others waiting: callers catch the error and fall back to TorchComm.
        def agree(ok: bool, what: str):
            flags = [None] * self.world
            self._dist.all_gather_object(flags, bool(ok), group=group)
            if not all(flags):
                try:
                    handle.peer_close()
                except Exception:  # noqa: BLE001
                    pass
                raise RuntimeError(f"peer-memory exchange unavailable: {what} failed on rank(s) "
                                   f"{[r for r, f in enumerate(flags) if not f]}")
        mine, err = None, None
        try:
            mine = handle.peer_create()
        except Exception as e:  # noqa: BLE001
            err = e
        handles = [None] * self.world
        self._dist.all_gather_object(handles, mine, group=group)
        agree(err is None and all(h_ is not None for h_ in handles), "b200mc_peer_create")
        try:
            handle.peer_connect(self.rank, self.world, handles)
        except Exception as e:  # noqa: BLE001
            err = e
        agree(err is None, "b200mc_peer_connect (cudaIpcOpenMemHandle)")
        self._dist.barrier(group=group)                     # every buffer is zeroed and mapped before the first store
        probe = self.allreduce_sum(np.array([1.0, float(self.rank)]))          # one real exchange as a self-test
        agree(bool(probe[0] == self.world and probe[1] == self.world * (self.world - 1) / 2), "the first exchange")

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        """Host arrays of up to 4352 doubles take the same exchange (copy in, one-shot all-reduce, copy out)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size == 0 or a.size > 4352:
            return super().allreduce_sum(a)
        flat = a.ravel()
        return self.allreduce_device(self.handle, flat.size, lambda ptr: self.handle.h2d(ptr, flat)).reshape(a.shape)

    def allreduce_device(self, handle, nelem: int, fill) -> np.ndarray:
        if handle is not self.handle or nelem > 4352:
            return super().allreduce_device(handle, nelem, fill)
        import torch
        dev = torch.device("cuda", handle.device)
        buf = getattr(self, "_buf", None)
        if buf is None or buf.numel() < nelem or buf.device != dev:
            buf = self._buf = torch.zeros(4352, dtype=torch.float64, device=dev)
        view = buf[:nelem]
        cur = torch.cuda.current_stream(dev)
        same = handle.stream == cur.cuda_stream
        if not same:
            cur.synchronize()
        if fill is None:
            view.zero_()
            if not same:
                cur.synchronize()
        else:
            fill(view.data_ptr())
        handle.peer_allreduce(view.data_ptr(), nelem)       # on the handle's stream, right behind the kernels
        if not same:
            handle.synchronize()
        return view.cpu().numpy()


def sharded_sums(handle, comm: Comm, params, spot, T, steps, n_paths, seed, strikes, is_call, flags,
                 bumps=None) -> np.ndarray:
    """Each rank simulates its path range; returns the all-reduced [n_strikes, NSUMS] sums on every rank."""
    lo, hi = shard_range(n_paths, comm.rank, comm.world)
    ks = np.atleast_1d(np.asarray(strikes, dtype=np.float64))
    if getattr(comm, "backend", None) == "nccl" and hasattr(comm, "allreduce_device") and hasattr(handle, "h"):
        fill = None
        if hi > lo:
            def fill(ptr):
                handle.price_european(params, spot, T, steps, hi - lo, seed, ks, is_call, flags, bumps, path_offset=lo,
                                      out_dev=ptr)
        return comm.allreduce_device(handle, ks.size * NSUMS, fill).reshape(ks.size, NSUMS)
    if hi > lo:
        local = handle.price_european(params, spot, T, steps, hi - lo, seed, ks, is_call, flags, bumps,
                                      path_offset=lo)
    else:
        local = np.zeros((ks.size, NSUMS))
    return comm.allreduce_sum(local).reshape(ks.size, NSUMS)


def sharded_generate_paths(handle, comm: Comm, params, spot, T, steps, n_paths, seed, flags=0, dtype=np.float32,
                           out_dev: Optional[int] = None):
    """Path-storing mode across ranks (BASELINE config 4): rank r simulates the contiguous range [lo, hi) of the GLOBAL
    path index into its OWN memory -- no exchange; rows are identical to those of a single-GPU run of n_paths paths.
    Returns (lo, hi, local) with local the [hi - lo, steps + 1] array (None when out_dev, a device pointer with room for
    (hi - lo) * (steps + 1) elements, is given)."""
    lo, hi = shard_range(n_paths, comm.rank, comm.world)
    if hi == lo:
        return lo, hi, (None if out_dev is not None else np.empty((0, steps + 1), dtype=dtype))
    local = handle.generate_paths(params, spot, T, steps, hi - lo, seed, flags, dtype, path_offset=lo, out_dev=out_dev)
    return lo, hi, local
