"""GPU: one handle shared by several Python threads (a web server's pool sharing the process-wide handle).  ctypes
releases the GIL during a call and the C side allows one call in flight per handle, so the binding serialises calls per
handle; results must equal the serial ones exactly."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_threads_sharing_one_handle_get_the_serial_results():
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    h = _lib.Handle(0)
    try:
        p, g = SVJParams(), SVJParams.gbm(0.3)
        jobs = [(p if i % 2 else g, 1000 + i, 20_000 + 997 * i, 50 + i) for i in range(24)]

        def run(job):
            prm, seed, n, steps = job
            row = h.price_european(prm, 2500.0, 0.5, steps, n, seed, [2400.0, 2500.0, 2600.0], True, _lib.FP64)
            S = h.simulate_terminal(prm, 2500.0, 0.5, steps, 257, seed, _lib.FP64, np.float64)[0]
            m = h.risk_metrics(S - 2500.0, 0.95)
            return row.copy(), S.copy(), m.copy()

        serial = [run(j) for j in jobs]
        out = [None] * len(jobs)
        errs = []

        def worker(k):
            try:
                for i in range(k, len(jobs), 6):
                    for _ in range(3):
                        out[i] = run(jobs[i])
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errs, errs
        for a, b in zip(out, serial):
            for x, y in zip(a, b):
                np.testing.assert_array_equal(x, y)
    finally:
        h.close()


@pytest.mark.parametrize("n", [257, 4001, 300_001, 4_000_000])
def test_tail_metrics_are_bitwise_reproducible(n):
    """The one-launch tail metrics gather their last candidates in the order of a global atomic counter; the candidates'
    tail sums are accumulated as exact integers so that this order cannot reach the result: repeated calls on the same
    vector (with another kernel's work in between) return identical bits."""
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    h = _lib.Handle(0)
    try:
        g = np.random.default_rng(n)
        for x in (g.standard_t(4, size=n) * 0.01, (g.standard_normal(n) * 37.0 - 5.0).astype(np.float32)):
            first = h.risk_metrics(x, 0.99)
            for i in range(12):
                h.price_european(SVJParams.gbm(0.3), 2500.0, 1.0, 50, 10_000 + 1000 * i, i, [2500.0])
                again = h.risk_metrics(x, 0.99)
                np.testing.assert_array_equal(again, first)
    finally:
        h.close()


def test_handles_do_not_leak_device_memory():
    """Create / use / destroy cycles: scratch, staging, pinned buffers, the tail-metric state and a pooled draws buffer are
    all released by b200mc_destroy."""
    import torch
    from monte_carlo_option_simulator_b200 import SVJParams, _lib

    def cycle(seed):
        h = _lib.Handle(0)
        try:
            p = SVJParams()
            h.price_european(p, 2500.0, 0.5, 60, 50_000, seed, [2400.0, 2500.0], True, _lib.GREEKS,
                             _lib.Bumps(0.01, p.v0 + 0.01, p.v0 - 0.01, p.r + 1e-4, p.r - 1e-4))
            S = h.simulate_terminal(p, 2500.0, 0.5, 60, 100_000, seed, 0, np.float64)[0]
            h.risk_metrics(S - 2500.0, 0.99)
            h.generate_paths(SVJParams.gbm(0.3), 2500.0, 1.0, 250, 20_000, seed, 0, np.float32)
            d = _lib.ReferenceDraws(h, seed, 5000, 50)
            d.simulate(p, 2500.0, 0.25)
            d.close()                                    # parks its buffer in the handle's one-slot pool
            h.numpy_fill(seed, 100_000)
        finally:
            h.close()

    cycle(0)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(25):
        cycle(i + 1)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 32 << 20, f"{(free0 - free1) >> 20} MiB lost over 25 handle cycles"


def test_two_handles_on_one_device_run_concurrently():
    """One Handle per thread is the way to overlap work on one device: each has its own stream and scratch.  Includes the
    cooperative tail-metric kernel on both at once."""
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    hs = [_lib.Handle(0), _lib.Handle(0)]
    try:
        p = SVJParams()
        x = np.random.default_rng(5).standard_t(4, size=1_000_001) * 0.01

        def run(h, seed):
            row = h.price_european(p, 2500.0, 0.5, 100, 400_000, seed, [2500.0], True, _lib.FP64)
            return row.copy(), h.risk_metrics(x, 0.99).copy()

        want = [run(hs[0], 1), run(hs[0], 2)]
        got = [[None] * 8, [None] * 8]
        errs = []

        def worker(k):
            try:
                for i in range(8):
                    got[k][i] = run(hs[k], k + 1)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=120)
        assert not errs and not any(t.is_alive() for t in ts), errs
        for k in range(2):
            for r in got[k]:
                np.testing.assert_array_equal(r[0], want[k][0])
                np.testing.assert_array_equal(r[1], want[k][1])
    finally:
        for h in hs:
            h.close()
