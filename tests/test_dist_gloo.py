"""CPU, world_size 2, gloo: the N > 1 path of the pricer -- contiguous shards of the GLOBAL path index, one all-reduce
of the per-strike sum vectors -- gives the single-process sums, and the engines built on it give the same dicts."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from test_host_logic import OracleBackedHandle
        from oracle import oracle as O
        from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, _lib
        from monte_carlo_option_simulator_b200.dist import TorchComm, shard_range, sharded_sums
        comm = TorchComm()
        assert (comm.rank, comm.world) == (rank, world)
        p = O.Params()
        h = OracleBackedHandle(n_global=1001)
        n = 1001                                   # odd: shards of 501 and 500 paths
        got = sharded_sums(h, comm, p, 22500.0, 0.25, 63, n, 5, [22000.0, 23000.0], True, _lib.ANTITHETIC)
        eng = MonteCarloEngine(p, n, 252, 5, use_sobol=False, rng="philox", handle=h, comm=comm)
        price = eng.price(22500.0, 22500.0, 0.25, True)
        delta = GreeksEngine(p, n, 252, 5, rng="philox", handle=h, comm=comm).delta(22500.0, 22500.0, 0.25, True)
        from test_host_logic import NumpyRiskHandle
        from monte_carlo_option_simulator_b200.risk import compute_risk_metrics_sharded
        x = np.random.default_rng(3).standard_t(3, size=20_001) * 0.01
        lo, hi = shard_range(x.size, rank, world)
        risk = compute_risk_metrics_sharded(x[lo:hi], 0.99, comm=comm, handle=NumpyRiskHandle())
        # scenario ladders over sharded paths: every cell's paths are split, the sums all-reduced once
        from monte_carlo_option_simulator_b200.risk import StressTestEngine
        many = eng.price_many([22500.0, 21000.0], [22500.0, 22000.0], 0.25, [True, False])
        stress = StressTestEngine(p, num_paths=n, seed=5, rng="philox", handle=h, comm=comm).jump_scenario(22500.0, 22500.0, 0.25)
        from monte_carlo_option_simulator_b200.risk import HedgingBacktest
        hedge = HedgingBacktest(p, seed=9, rng="philox", handle=OracleBackedHandle(), comm=comm).run_backtest(
            22500.0, 22500.0, 0.05, True, num_scenarios=11, num_mc_paths=200)       # scenarios 0..5 | 6..10
        q.put((rank, got, price, delta, shard_range(n, rank, world), risk, many, stress, hedge))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_process():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_logic import OracleBackedHandle
    from oracle import oracle as O
    from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, _lib

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0

    p = O.Params()
    h = OracleBackedHandle()
    whole = h.price_european(p, 22500.0, 0.25, 63, 1001, 5, [22000.0, 23000.0], True, _lib.ANTITHETIC)
    price = MonteCarloEngine(p, 1001, 252, 5, use_sobol=False, rng="philox", handle=h).price(22500.0, 22500.0, 0.25, True)
    delta = GreeksEngine(p, 1001, 252, 5, rng="philox", handle=h).delta(22500.0, 22500.0, 0.25, True)
    assert results[0][4] == (0, 501) and results[1][4] == (501, 1001)
    want_risk = O.risk_metrics(np.random.default_rng(3).standard_t(3, size=20_001) * 0.01, 0.99)
    eng1 = MonteCarloEngine(p, 1001, 252, 5, use_sobol=False, rng="philox", handle=h)
    want_many = [eng1.price(22500.0, 22500.0, 0.25, True), eng1.price(21000.0, 22000.0, 0.25, False)]
    from monte_carlo_option_simulator_b200.risk import StressTestEngine
    want_stress = StressTestEngine(p, num_paths=1001, seed=5, rng="philox", handle=h).jump_scenario(22500.0, 22500.0, 0.25)
    from monte_carlo_option_simulator_b200.risk import HedgingBacktest
    from conftest import assert_tree_close
    want_hedge = HedgingBacktest(p, seed=9, rng="philox", handle=OracleBackedHandle()).run_backtest(
        22500.0, 22500.0, 0.05, True, num_scenarios=11, num_mc_paths=200)
    for rank, got, pr_, dl, _, risk, many, stress, hedge in results:
        assert_tree_close(hedge, want_hedge, rel=1e-10, abs_=1e-9)
        for g_, w_ in zip(many, want_many):
            assert g_ == pytest.approx(w_, rel=1e-10, abs=1e-9)
        assert stress == pytest.approx(want_stress, rel=1e-9, abs=1e-8)
        for k, w in want_risk.items():
            assert risk[k] == pytest.approx(w, rel=1e-10, abs=1e-13), k
        np.testing.assert_allclose(got, whole, rtol=1e-12)
        for k, w in price.items():
            assert pr_[k] == pytest.approx(w, rel=1e-10, abs=1e-9)
        for k, w in delta.items():
            assert dl[k] == pytest.approx(w, rel=1e-9, abs=1e-9)
