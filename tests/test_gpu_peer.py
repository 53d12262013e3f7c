"""GPU: the single-rank leg of the peer-memory exchange (csrc/peer.cu).  With one rank the all-reduce is the identity and
exercises buffer creation, the epoch / parity logic and the argument checks; the multi-rank behaviour (equal to NCCL,
bitwise identical across ranks, 300 epochs) is checked by tools/peer_allreduce_check.py under torchrun
(profiles/r01_peer_allreduce_n2.txt) because CUDA IPC needs one process per GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_rank_exchange_is_the_identity_and_checks_arguments():
    from monte_carlo_option_simulator_b200 import _lib
    h = _lib.Handle(0)
    try:
        buf = h.malloc(4352 * 8)
        with pytest.raises(_lib.B200MCError, match="connect"):
            h.peer_allreduce(buf, 17)
        ipc = h.peer_create()
        assert len(ipc) == 64 and h.peer_create() == ipc            # idempotent: the same buffer
        with pytest.raises(_lib.B200MCError):
            h.peer_connect(0, 2, [ipc])                             # one handle per rank
        with pytest.raises(_lib.B200MCError):
            h.peer_connect(3, 1, [ipc])
        h.peer_connect(0, 1, [ipc])
        g = np.random.default_rng(0)
        for n in (1, 17, 136, 4352, 17, 17):                        # several epochs, both parities
            x = g.standard_normal(n)
            h.h2d(buf, x)
            h.peer_allreduce(buf, n)
            y = np.empty(n)
            h.d2h(y, buf)
            np.testing.assert_array_equal(x, y)
        with pytest.raises(_lib.B200MCError):
            h.peer_allreduce(buf, 4353)
        # the fused kernel's device-resident sums go through the exchange unchanged
        from monte_carlo_option_simulator_b200 import SVJParams
        p = SVJParams.gbm(0.3, r=0.065)
        want = h.price_european(p, 2500.0, 1.0, 50, 10_000, 3, [2500.0], True, 0)
        h.price_european(p, 2500.0, 1.0, 50, 10_000, 3, [2500.0], True, 0, out_dev=buf)
        h.peer_allreduce(buf, _lib.NSUMS)
        got = np.empty((1, _lib.NSUMS))
        h.d2h(got, buf)
        np.testing.assert_array_equal(got, want)
        # sharded tail metrics with a single shard == the single-device metrics (multi-kernel select + the exchange in
        # between vs the one-launch select: same order statistics, sums folded in a different order)
        x = g.standard_t(3, size=200_001) * 0.01
        for conf in (0.99, 0.5):
            a, b = h.risk_metrics_sharded(x, conf), h.risk_metrics(x, conf)
            assert a[0] == b[0]                                        # VaR: an order statistic, exact
            np.testing.assert_allclose(a, b, rtol=1e-11, equal_nan=True)
        x32 = x[:5000].astype(np.float32)
        np.testing.assert_allclose(h.risk_metrics_sharded(x32), h.risk_metrics(x32), rtol=1e-11, equal_nan=True)
        with pytest.raises(_lib.B200MCError, match="empty"):
            h.risk_metrics_sharded(np.zeros(0))
        h.peer_close()
        with pytest.raises(_lib.B200MCError):
            h.peer_allreduce(buf, 17)
        h.free(buf)
    finally:
        h.close()
