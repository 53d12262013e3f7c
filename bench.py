#!/usr/bin/env python
"""bench.py -- GBM path-steps/s of the fused European kernel (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 2 -- European call + the Delta/Gamma/Vega/Rho bump accumulators in ONE
fused launch, S0 = K = 2500, r = 6.5 %, sigma = 30 %, T = 1 y, 250 steps, 10M paths per GPU (weak scaling: rank g
simulates global paths [g*1e7, (g+1)*1e7) of the same Philox key), fp32 path state, fp64 sums.  A "step" is one
such pricing pass; for N > 1 it ends with the path's only exchange, an NCCL all-reduce of the 17-double sum vector.

value  = N * paths * 250 * K / t, t = CUDA-event time of the K steps on the launching stream, max over ranks.
e2e    = the same metric through the public Python API (GreeksEngine.delta/vega/gamma -> one C-ABI call with host
         arguments and a host result per step: strikes go host->device, the b200mc_sums struct device->host).
roofline = instruction roofline of the fused kernel (it moves no data): per-path-step instruction counts of the
         kernel's hot loop (read from the SASS of the shipped .so) against issue rates of the same pipes measured
         in this run by b200mc_microbench.  roofline_hbm = the path-store kernel against MEASURED_PEAKS.json.
cpu_baseline = the oracle's port of the reference CPU path (NumPy PCG64 draws + OpenMP C recurrence + NumPy
         reduction) on this box's host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gbm_path_steps_per_sec"
UNIT = "path-steps/s"
PATHS_PER_GPU = 10_000_000
N_STEPS = 250
SPOT = STRIKE = 2500.0
T = 1.0
WORKLOAD = ("BASELINE cfg2: European call + Delta/Gamma/Vega/Rho accumulators, one fused launch, "
            "S0=K=2500 r=6.5% sigma=30% T=1y, 250 steps, 10M paths per GPU")

# fallback instruction mix per Philox call (8 path-steps) of k_european<GBM, no anti, greeks, fp32> (tools/sass_mix.py)
FALLBACK_MIX = {"heavy": 18, "alu": 33, "fp32": 20, "xu": 16, "uni": 2, "lsu": 2, "ctl": 1, "total": 92, "imad_wide": 17, "philox_calls": 1}


def gbm_params():
    from monte_carlo_option_simulator_b200 import SVJParams
    return SVJParams.gbm(0.30, r=0.065, q=0.0)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """Samples before this call (pre-roll, GPU idle) are not reported."""
        self.first = len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(steps, warmup, sample_paths=50_000):
    """The reference's CPU path for this workload, restated by the oracle: per step one plain-MC price (RNG front
    end + recurrence + reduction, monte_carlo.py:273-375) and the delta / vega / gamma re-simulations
    (greeks.py:53-203) on `sample_paths` x 250, all host threads OpenMP can use.  Returns (path-steps/s, cores, s/step)."""
    from oracle import oracle as O
    O.build()
    p = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.0, r=0.065, q=0.0)
    eng = O.MonteCarloOracle(p, sample_paths, N_STEPS, 42, use_sobol=False, use_antithetic=False, use_control_variate=False)
    grk = O.GreeksOracle(p, sample_paths, N_STEPS, 42)

    def one():
        eng.price(SPOT, STRIKE, T, True)
        grk.delta(SPOT, STRIKE, T, True)
        grk.vega(SPOT, STRIKE, T, True)
        grk.gamma(SPOT, STRIKE, T, True)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_paths * N_STEPS / dt, O.num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is entitled to every host thread
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    sample = 50_000
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    v, cores, dt = cpu_reference_steps(steps, warm, sample)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{sample} paths x {N_STEPS} steps per step"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} paths x {N_STEPS} steps: price + delta + vega + gamma per step "
                                       "(NumPy PCG64 draws, OpenMP C recurrence, NumPy reductions)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ roofline
def instruction_roofline(h, achieved_path_steps_per_gpu, sm_mhz):
    """Instruction roofline of k_european<GBM, fp32, greeks>: the hot loop's per-path-step instruction counts by
    pipe (SASS of the shipped library) over the issue rates measured now on this device."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_mix
        mixes = sass_mix.mix("k_europeanILi0ELb0ELb1EfLb1E")
        mix = list(mixes.values())[0]
        src = "cuobjdump -sass of libb200mc.so (tools/sass_mix.py)"
        if not mix.get("xu"):
            raise RuntimeError("empty mix")
    except Exception as e:  # noqa: BLE001
        mix, src = dict(FALLBACK_MIX), f"fallback constants ({type(e).__name__})"
    steps_per_iter = 8.0 * mix.get("philox_calls", 1)
    per_step = {k: mix.get(k, 0) / steps_per_iter for k in ("heavy", "alu", "fp32", "xu", "uni", "lsu", "ctl", "total", "imad_wide")}
    rates = {"ffma": h.microbench(0), "imad_wide": h.microbench(1), "lop3": h.microbench(2),
             "mufu_ex2": h.microbench(3), "mufu_sin": h.microbench(4), "mufu_lg2": h.microbench(9),
             "mufu_sqrt": h.microbench(10), "ffma_lop3_pairs": h.microbench(11),
             "philox_calls": h.microbench(6), "philox_bm_calls": h.microbench(7)}
    xu = min(rates["mufu_ex2"], rates["mufu_sin"], rates["mufu_lg2"], rates["mufu_sqrt"])

    def bounds_for(n_wide, n_other, n_alu, n_xu):
        # serial-issue model measured by tools/pipe_probe.py: an IMAD.WIDE holds the sub-partition's issue port for
        # 1/R_wide (4 cycles), every other instruction for one issue slot (1/R_ffma); XU and ALU work overlaps.
        return {"issue": 1.0 / (n_wide / rates["imad_wide"] + n_other / rates["ffma"]),
                "xu": xu / n_xu if n_xu else float("inf"),
                "alu": rates["lop3"] / n_alu if n_alu else float("inf")}

    # ALGORITHMIC cost per path-step: one Philox4x32-10 call per 8 steps with the first round's
    # multiplies loop-invariant/uniform = 17 IMAD.WIDE + 20 LOP3; per Box-Muller pair 2 ALU + 3 FP32 + 4 MUFU + 2 FP32 to
    # scale and accumulate; ~4 loop instructions per call.
    algo = {"imad_wide": 17 / 8, "alu": (20 + 4 * 2) / 8, "fp32": 4 * 5 / 8, "xu": 2.0, "loop": 4 / 8}
    algo_other = algo["alu"] + algo["fp32"] + algo["xu"] + algo["loop"]
    b_algo = bounds_for(algo["imad_wide"], algo_other, algo["alu"], algo["xu"])
    b_sass = bounds_for(per_step["imad_wide"], per_step["total"] - per_step["imad_wide"], per_step["alu"], per_step["xu"])
    binding = min(b_algo, key=b_algo.get)
    peak = b_algo[binding]
    return {"bound": binding, "achieved": achieved_path_steps_per_gpu, "peak": peak, "unit": UNIT,
            "frac": achieved_path_steps_per_gpu / peak, "traffic": 23040,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture "
                              "profiles/r01_ncu_k_european_gbm_fp32_greeks.txt (algorithmic bytes: 0 per path-step)",
            "kernel": "k_european<GBM, fp32, greeks>", "kind": "instruction roofline (the kernel moves no data): "
            "algorithmic instructions per path-step over issue rates measured in this run",
            "algorithmic_per_path_step": algo, "pipe_bounds_algorithmic": b_algo,
            "sass_per_path_step": per_step, "pipe_bounds_sass_mix": b_sass,
            "frac_of_sass_mix_bound": achieved_path_steps_per_gpu / min(b_sass.values()), "mix_source": src,
            "measured_rates_ops_per_s": rates,
            "peak_source": "b200mc_microbench on this device in this run (not in MEASURED_PEAKS.json)",
            "rng_only_path_steps_per_s": 8.0 * rates["philox_bm_calls"]}


def hbm_roofline(h, torch, n_paths=4_000_000, reps=3):
    """BASELINE cfg4: the path-store kernel (4M paths x 250 steps; 4 or 8 bytes per path-step, written once, output
    far larger than L2) vs the measured copy bandwidth."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, read+write)"
    else:
        peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    from monte_carlo_option_simulator_b200 import _lib
    out = {}
    cases = (("f32", np.float32, 0, 4, N_STEPS + 1, "reference layout [n, 251] float32: CTA tile = 32 paths x 251, one TMA bulk store per tile"),
             ("f64_out_f32_state", np.float64, 0, 8, N_STEPS + 1, "reference layout [n, 251] float64 (what get_sample_paths returns), fp32 path state"),
             ("f64", np.float64, _lib.FP64, 8, N_STEPS + 1, "reference layout [n, 251] float64, fp64 path state (double exp per step: FP64-pipe bound)"),
             ("f32_ld256", np.float32, 0, 4, 256, "rows padded to 256 elements: row-tiled kernel, 128-bit shared + global stores"))
    for name, dt_np, fl, esz, ld, note in cases:
        buf = torch.empty(n_paths * ld * esz, dtype=torch.uint8, device="cuda")
        best = None
        for r in range(reps + 1):
            h.timer_begin()
            h.generate_paths(gbm_params(), SPOT, T, N_STEPS, n_paths, 42 + r, fl, dt_np, 0, ld, out_dev=buf.data_ptr())
            ms = h.timer_end()
            if r > 0:
                best = ms if best is None else min(best, ms)
        nbytes = n_paths * (N_STEPS + 1) * esz          # algorithmic bytes: the matrix itself, no padding
        gbs = nbytes / (best * 1e-3) / 1e9
        out[name] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": 3956231456 if name == "f32" else None,
                     "traffic_source": "ncu --set full, profiles/r01_ncu_k_paths_tma_fp32.txt (4M x 251 fp32 launch)" if name == "f32" else None,
                     "kernel": "k_paths_tma<GBM>" if ld == N_STEPS + 1 else "k_paths_det<GBM>", "path_steps_per_s": n_paths * N_STEPS / (best * 1e-3),
                     "bytes_per_launch": nbytes, "ms": best, "peak_source": src, "shape": [n_paths, N_STEPS + 1], "ld": ld,
                     "note": note}
        del buf
    return out


# ------------------------------------------------------------------------------------------------ own arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from monte_carlo_option_simulator_b200 import GreeksEngine, _lib
    from monte_carlo_option_simulator_b200.dist import TorchComm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = _lib.Handle(local)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    p = gbm_params()
    n = args.paths
    bumps = _lib.Bumps(0.01, p.v0 + 0.01, max(p.v0 - 0.01, 0.001), p.r + 1e-4, max(p.r - 1e-4, 0))
    flags = _lib.GREEKS
    out = torch.zeros(_lib.NSUMS, dtype=torch.float64, device="cuda")

    # the exchange step (17 fp64 sums per step): one-shot all-reduce over NVLink peer memory (csrc/peer.cu) unless
    # B200MC_EXCHANGE=nccl asks for the NCCL call
    use_peer = world > 1 and os.environ.get("B200MC_EXCHANGE", "peer") != "nccl"
    comm = None
    if world > 1:
        from monte_carlo_option_simulator_b200.dist import PeerComm
        if use_peer:
            try:
                comm = PeerComm(h)              # raises on ALL ranks when CUDA IPC / P2P is not available on any of them
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] {e}; using the NCCL all-reduce", file=sys.stderr, flush=True)
                use_peer = False
        if not use_peer:
            comm = TorchComm()

    def step(i):
        h.price_european(p, SPOT, T, N_STEPS, n, 42 + i, [STRIKE], True, flags, bumps, path_offset=rank * n,
                         out_dev=out.data_ptr())
        if use_peer:
            h.peer_allreduce(out.data_ptr(), _lib.NSUMS)
        elif world > 1:
            dist.all_reduce(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launches - l0
    # The timed region is tens of milliseconds, shorter than nvidia-smi's sampling period, so the sampler keeps
    # running over ~1 s of the SAME steps issued back to back right after it (not timed): the reported clocks, power
    # and throttle reasons are those of this workload under sustained load.
    t_end = time.perf_counter() + 1.0
    k = 0
    while time.perf_counter() < t_end:
        for _ in range(8):
            step(args.warmup + args.steps + k)
            k += 1
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "timed region + 1 s of the same steps back to back"
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * N_STEPS * args.steps / (ms * 1e-3)
    sums = out.cpu().numpy()

    # ---- e2e through the public API (host arguments in, host dict out) -----------------------------------
    g = GreeksEngine(p, n * world, N_STEPS, seed=1000, rng="philox", handle=h, comm=comm)
    e2e_steps = args.steps

    def api_step(i):
        g.seed = 1000 + i
        return g.delta(SPOT, STRIKE, T, True), g.vega(SPOT, STRIKE, T, True), g.gamma(SPOT, STRIKE, T, True)

    api_step(-1)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        last = api_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * N_STEPS * e2e_steps / float(t.item())

    # BASELINE cfg3, benchmark reading: 64 strikes x 16 expiries, 1M paths PER CELL (disjoint counter ranges), cells dealt
    # round-robin to the ranks, no path-level collective (one all-reduce gathers the 1024 sum vectors)
    from monte_carlo_option_simulator_b200 import MonteCarloEngine
    cfg3 = None
    if not args.no_extras:
        eng3 = MonteCarloEngine(p, 1_000_000, N_STEPS, 42, use_sobol=False, use_antithetic=False, use_control_variate=False,
                                rng="philox", handle=h, comm=comm)
        ks3, Ts3 = np.linspace(0.7, 1.3, 64) * SPOT, [j / 8 for j in range(1, 17)]
        eng3.price_grid(SPOT, ks3[:2], Ts3[:2], True, independent_cells=True)          # warm-up (kernel load)
        barrier()
        t0 = time.perf_counter()
        g3 = eng3.price_grid(SPOT, ks3, Ts3, True, independent_cells=True)
        barrier()
        dt3 = time.perf_counter() - t0
        work3 = 64 * 1_000_000 * float(g3["num_steps"].sum())
        cfg3 = {"seconds": dt3, "cells": 1024, "path_steps": work3, "path_steps_per_s": work3 / dt3,
                "reading": "independent cells: 1M paths per (expiry, strike) cell, cells sharded over the ranks",
                "atm_1y_price": float(g3["prices"][7, 32]), "atm_1y_std_error": float(g3["std_errors"][7, 32])}

    if rank == 0:
        disc = float(np.exp(-p.r * T))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "paths_per_gpu": n, "n_steps": N_STEPS, "rng": "Philox4x32-10 in registers",
                           "exchange": ("none" if world == 1 else
                                        "one-shot all-reduce of 17 fp64 sums per step over NVLink peer memory (k_peer_allreduce)"
                                        if use_peer else "NCCL all-reduce of 17 fp64 sums per step"),
                           "l2": "kernel reads no global inputs (counter-based RNG), nothing to flush; every step uses a new seed"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": _lib.NSUMS * 8,
                        "api": "GreeksEngine.delta + .vega + .gamma (one fused launch behind them), host dicts out",
                        "ms_per_step": e2e_s / e2e_steps * 1e3},
                "cfg3_grid_independent_cells": cfg3,
                "check": {"price": disc * sums[1] / sums[0], "delta_pathwise": disc * sums[9] / sums[0],
                          "api_delta": last[0]["pathwise"], "api_gamma": last[2]["gamma"]}}
        if world == 1 and not args.no_extras:
            sm_mhz = clocks.get("sm_mhz") if clocks else None
            line["roofline"] = instruction_roofline(h, value, sm_mhz)
            try:
                line["roofline_hbm"] = hbm_roofline(h, torch)
            except Exception as e:  # noqa: BLE001
                line["roofline_hbm"] = {"error": str(e)}
            extras = {}
            from monte_carlo_option_simulator_b200 import SVJParams

            def rate(pp, spot, npaths, fl, ks):
                best = None
                for r in range(4):                                  # first call loads the kernel: not timed
                    h.timer_begin()
                    h.price_european(pp, spot, T, N_STEPS, npaths, 7 + r, ks, True, fl, None, out_dev=out_many.data_ptr())
                    ms_ = h.timer_end()
                    if r:
                        best = ms_ if best is None else min(best, ms_)
                return npaths * N_STEPS / (best * 1e-3)

            out_many = torch.zeros(64 * _lib.NSUMS, dtype=torch.float64, device="cuda")
            extras["fp32_price_only"] = rate(p, SPOT, n, 0, [STRIKE])
            extras["fp32_antithetic"] = rate(p, SPOT, n, _lib.ANTITHETIC, [STRIKE])
            extras["fp64_price_only"] = rate(p, SPOT, n, _lib.FP64, [STRIKE])
            extras["fp32_64_strikes_shared_paths"] = rate(p, SPOT, n, _lib.ANTITHETIC, list(np.linspace(0.7, 1.3, 64) * SPOT))
            extras["fp32_heston_antithetic"] = rate(SVJParams(lambda_j=0.0), 22500.0, n // 4, _lib.ANTITHETIC, [22500.0])
            extras["fp32_svj_antithetic"] = rate(SVJParams(), 22500.0, n // 4, _lib.ANTITHETIC, [22500.0])
            # BASELINE cfg3, API-parity reading: 64 strikes x 16 expiries, 1M paths per expiry shared across strikes
            from monte_carlo_option_simulator_b200 import MonteCarloEngine
            eng = MonteCarloEngine(p, 1_000_000, N_STEPS, 42, use_sobol=False, use_antithetic=False,
                                   use_control_variate=False, rng="philox", handle=h)
            ks, Ts = np.linspace(0.7, 1.3, 64) * SPOT, [j / 8 for j in range(1, 17)]
            eng.price_grid(SPOT, ks, Ts, True)
            t0 = time.perf_counter()
            grid = eng.price_grid(SPOT, ks, Ts, True)
            dtg = time.perf_counter() - t0
            extras["cfg3_grid_64x16_1M_paths"] = 1_000_000 * float(grid["num_steps"].sum()) / dtg
            line["cfg3_grid"] = {"seconds": dtg, "cells": 1024, "path_steps": 1_000_000 * int(grid["num_steps"].sum()),
                                 "reading": "price_batch semantics: paths shared across the 64 strikes of an expiry",
                                 "atm_1y_price": float(grid["prices"][7, 32])}
            line["extras_path_steps_per_s"] = extras
            # the callers of SURVEY 8(f): scenario grids, hedging backtest, implied-vol chain, quasi-Monte Carlo
            try:
                from monte_carlo_option_simulator_b200.risk import HedgingBacktest, StressTestEngine
                nxt = {}
                cells = _lib.make_cells(p, 22500.0, 0.25, 63, 50_000, 42 + np.arange(1000))
                outc = torch.zeros(1000 * _lib.NSUMS, dtype=torch.float64, device="cuda")
                best = None
                for r in range(4):
                    h.timer_begin()
                    h.price_cells(cells, np.full(1000, 22500.0), _lib.ANTITHETIC, out_dev=outc.data_ptr())
                    ms_ = h.timer_end()
                    if r:
                        best = ms_ if best is None else min(best, ms_)
                nxt["cells_1000x50k_paths_x63_steps"] = {"ms": best, "path_steps_per_s": 1000 * 50_000 * 63 / (best * 1e-3)}
                st = StressTestEngine(SVJParams(), num_paths=200_000, seed=42, handle=h)
                bt = HedgingBacktest(SVJParams(), seed=42, handle=h)
                st.full_stress_report(22500.0, 22500.0, 0.25)
                bt.run_backtest(22500.0, 22500.0, 0.25, num_scenarios=10, num_mc_paths=1000)
                t0 = time.perf_counter()
                st.full_stress_report(22500.0, 22500.0, 0.25)
                nxt["svj_full_stress_report_200k_paths_s"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                bt.run_backtest(22500.0, 22500.0, 0.25)
                nxt["svj_hedging_backtest_1000x50k_paths_s"] = time.perf_counter() - t0
                g = np.random.default_rng(1)
                kk, tt = 22500.0 * g.uniform(0.8, 1.2, 2048), g.uniform(0.05, 1.0, 2048)
                h.implied_vol(np.full(2048, 900.0), 22500.0, kk, tt, 0.065, 0.012)
                t0 = time.perf_counter()
                h.implied_vol(np.full(2048, 900.0), 22500.0, kk, tt, 0.065, 0.012)
                nxt["implied_vol_2048_options_s"] = time.perf_counter() - t0
                qe = MonteCarloEngine(p, 1 << 20, N_STEPS, 42, use_antithetic=False, use_control_variate=False, rng="sobol",
                                      handle=h)
                qe.price(SPOT, STRIKE, T)
                t0 = time.perf_counter()
                qp = qe.price(SPOT, STRIKE, T)
                from monte_carlo_option_simulator_b200 import bs_price
                nxt["qmc_1M_paths_x250"] = {"seconds": time.perf_counter() - t0,
                                            "abs_error_vs_black_scholes": abs(qp["price"] - bs_price(SPOT, STRIKE, T, p.r, p.q, 0.3, True))}
                line["next_rows"] = nxt
            except Exception as e:  # noqa: BLE001
                line["next_rows"] = {"error": str(e)}
            v, cores, dt = cpu_reference_steps(2, 1, 50_000)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "50000 paths x 250 steps: price + delta + vega + gamma per step, 2 steps "
                                              "(NumPy PCG64 draws, OpenMP C recurrence, NumPy reductions)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths", type=int, default=PATHS_PER_GPU, help="paths per GPU and step")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
