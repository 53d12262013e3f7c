#!/usr/bin/env python
"""bench.py -- GBM path-steps/s of the fused European kernel (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 2 -- European call + the Delta/Gamma/Vega/Rho bump accumulators in ONE
fused launch, S0 = K = 2500, r = 6.5 %, sigma = 30 %, T = 1 y, 250 steps, 10M paths per GPU (weak scaling: rank g
simulates global paths [g*1e7, (g+1)*1e7) of the same Philox key), fp32 path state, fp64 sums.  A "step" is one
such pricing pass; for N > 1 it ends with the path's only exchange, the all-reduce of the 17-double sum vector
(k_peer_allreduce over NVLink peer memory; NCCL with B200MC_EXCHANGE=nccl).

value  = N * paths * 250 * K / t, t = CUDA-event time of the K steps on the launching stream, max over ranks.
e2e    = the same metric through the public Python API (GreeksEngine.delta/vega/gamma -> one C-ABI call with host
         arguments and a host result per step: strikes go host->device, the b200mc_sums struct device->host).
roofline = instruction roofline of the fused kernel (it moves no data): per-path-step instruction counts of the
         kernel's hot loop (read from the SASS of the shipped .so) against issue rates of the same pipes measured
         in this run by the probe library (tools/probe).  roofline_hbm = the path-store kernel against MEASURED_PEAKS.json.
roofline_fp64 / roofline_heston / roofline_svj = the same model for the fp64 leg of cfg2 and for the reference's default
         model (Heston + jumps) and its jump-free special case.
cfg5_strong_scaling, cfg4_sharded_path_store, cfg2_fp64_greeks, check.allreduce = the multi-GPU configurations of
         BASELINE.json measured inside this run at every N (CUDA events, max over ranks).
cpu_baseline = the reference ITSELF (baseline/_ref, staged by tools/install_reference.py; kind "reference") -- or, when
         that copy is absent, the oracle's port (kind "port") -- on this box's host cores, on a bounded sample, with
         BASELINE.md section 3's two readings (kernel only, price() end to end).
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")          # tools/install_reference.py (git-ignored copy of the reference)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ReferenceCPU:
    """The reference's own CPU implementation of the workload, imported UNMODIFIED from baseline/_ref (kind
    "reference"): MonteCarloEngine.price (engine/monte_carlo.py:273-375) + GreeksEngine.delta/.vega/.gamma
    (engine/greeks.py:53-203) on the Numba kernel _simulate_svj_paths_numba (:189-243), all Numba threads."""
    kind = "reference"

    def __init__(self, sample_paths):
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_b200mc")
        os.environ["NUMBA_NUM_THREADS"] = str(_host_threads())       # torchrun pins OMP_NUM_THREADS=1 for its workers
        os.environ["OMP_NUM_THREADS"] = str(_host_threads())
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import numba
        from engine.greeks import GreeksEngine
        from engine.models import SVJParams
        from engine.monte_carlo import MonteCarloEngine, _simulate_svj_paths_numba
        self.n = sample_paths
        self.threads = numba.get_num_threads()
        self.kernel = _simulate_svj_paths_numba
        # BASELINE.md section 3 parameters (the GBM special case of the reference's SVJ engine)
        self.p = SVJParams(kappa=3.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0)
        self.eng = MonteCarloEngine(self.p, sample_paths, N_STEPS, seed=42, use_sobol=False, use_antithetic=False,
                                    use_control_variate=False)
        self.grk = GreeksEngine(self.p, sample_paths, N_STEPS, seed=42)
        self.what = ("reference MonteCarloEngine.price + GreeksEngine.delta/.vega/.gamma (NumPy PCG64 draws, Numba prange kernel, "
                     f"{self.threads} Numba threads)")

    def step(self):
        a = self.eng.price(SPOT, STRIKE, T, True)
        self.grk.delta(SPOT, STRIKE, T, True)
        self.grk.vega(SPOT, STRIKE, T, True)
        self.grk.gamma(SPOT, STRIKE, T, True)
        return a

    def readings(self, best_of=5):
        """(i) the kernel alone on pre-generated float64 draws, (ii) price() end to end."""
        p, n = self.p, self.n
        g = np.random.default_rng(42)
        Z = [g.standard_normal((n, N_STEPS)) for _ in range(3)]
        U = np.random.default_rng(43).random((n, N_STEPS))
        args = (float(SPOT), p.v0, p.r, p.q, T, p.kappa, p.theta, p.xi, p.rho, p.lambda_j, p.mu_j, p.sigma_j, Z[0], Z[1], U, Z[2], N_STEPS)
        self.kernel(*args)
        tk = min(_timed(lambda: self.kernel(*args)) for _ in range(best_of))
        self.eng.price(SPOT, STRIKE, T, True)
        tp = min(_timed(lambda: self.eng.price(SPOT, STRIKE, T, True)) for _ in range(max(2, best_of // 2)))
        return {"kernel_only_path_steps_per_s": n * N_STEPS / tk, "kernel_only_ms": tk * 1e3,
                "price_end_to_end_path_steps_per_s": n * N_STEPS / tp, "price_end_to_end_ms": tp * 1e3,
                "what": "best-of-N wall time at 50k x 250 (BASELINE.md section 3): (i) _simulate_svj_paths_numba on pre-generated "
                        "float64 draws, (ii) MonteCarloEngine.price(sobol/antithetic/CV off) incl. its NumPy RNG front end"}


class PortCPU:
    """Fallback when baseline/_ref is absent: the oracle's port of the same calls (kind "port")."""
    kind = "port"

    def __init__(self, sample_paths):
        os.environ["OMP_NUM_THREADS"] = str(_host_threads())
        from oracle import oracle as O
        O.build()
        p = O.Params(kappa=0.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.0, r=0.065, q=0.0)
        self.n = sample_paths
        self.eng = O.MonteCarloOracle(p, sample_paths, N_STEPS, 42, use_sobol=False, use_antithetic=False, use_control_variate=False)
        self.grk = O.GreeksOracle(p, sample_paths, N_STEPS, 42)
        self.threads = O.num_threads()
        self.what = f"oracle port: NumPy PCG64 draws, OpenMP C recurrence ({self.threads} threads), NumPy reductions"

    def step(self):
        a = self.eng.price(SPOT, STRIKE, T, True)
        self.grk.delta(SPOT, STRIKE, T, True)
        self.grk.vega(SPOT, STRIKE, T, True)
        self.grk.gamma(SPOT, STRIKE, T, True)
        return a

    def readings(self, best_of=3):
        self.eng.price(SPOT, STRIKE, T, True)
        tp = min(_timed(lambda: self.eng.price(SPOT, STRIKE, T, True)) for _ in range(best_of))
        return {"price_end_to_end_path_steps_per_s": self.n * N_STEPS / tp, "price_end_to_end_ms": tp * 1e3}


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def make_cpu_arm(sample_paths=50_000):
    if os.path.isdir(os.path.join(REF_DIR, "engine")):
        try:
            return ReferenceCPU(sample_paths)
        except Exception as e:  # noqa: BLE001  (numba missing, ...): say so and time the port
            print(f"[bench] reference in baseline/_ref not usable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    return PortCPU(sample_paths)


def cpu_baseline_block(arm, value, dt, steps, readings=None):
    """One workload step = price + delta + vega + gamma at sample x 250, i.e. what ONE fused GPU launch returns; the
    reference simulates 10 path sets and regenerates its draws 4 times for it, so per SIMULATED
    path-step its rate is 10x the workload-level `value` -- both are stated."""
    b = {"value": value, "unit": UNIT, "cores": arm.threads, "kind": arm.kind, "cpu_model": cpu_model(),
         "host_threads_available": _host_threads(),
         "sample": f"{arm.n} paths x {N_STEPS} steps: price + delta + vega + gamma per step, {steps} step(s); {arm.what}",
         "ms_per_step": dt * 1e3, "simulated_path_steps_per_s": 10.0 * value,
         "note": "value counts one 50k x 250 unit per step although the CPU arm simulates 10 path sets for it"}
    if readings:
        b["baseline_md_section3"] = readings
    return b


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = make_cpu_arm(50_000)
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 5))
    for _ in range(warm):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step()
    dt = (time.perf_counter() - t0) / steps
    v = arm.n * N_STEPS / dt
    try:
        readings = arm.readings()
    except Exception as e:  # noqa: BLE001
        readings = {"error": str(e)}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{arm.n} paths x {N_STEPS} steps per step (BASELINE cfg1 size)"},
            "cpu_baseline": cpu_baseline_block(arm, v, dt, steps, readings),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ roofline
def measured_rates(h):
    """Issue rates of the pipes the fused kernels live on, measured on this device in this run by the probe library
    (tools/probe/libb200mc_probe.so -- a measurement tool with its own C ABI, not part of libb200mc.so)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from probe import probe as P
    dev = h.device
    r = {"ffma": P.rate(P.FFMA, device=dev), "imad_wide": P.rate(P.IMAD_WIDE, device=dev), "lop3": P.rate(P.LOP3, device=dev),
         "mufu_ex2": P.rate(P.MUFU_EX2, device=dev), "mufu_sin": P.rate(P.MUFU_SIN, device=dev),
         "mufu_lg2": P.rate(P.MUFU_LG2, device=dev), "mufu_sqrt": P.rate(P.MUFU_SQRT, device=dev),
         "ffma_lop3_pairs": P.rate(P.FFMA_LOP3, device=dev), "f2f_f32_f64": P.rate(P.F2F, device=dev),
         "dadd": P.rate(P.DADD, device=dev), "philox_calls": P.rate(P.PHILOX, device=dev),
         "philox_bm_calls": P.rate(P.PHILOX_BM, device=dev)}
    r["xu"] = min(r["mufu_ex2"], r["mufu_sin"], r["mufu_lg2"], r["mufu_sqrt"])
    return r


def pipe_bounds(rates, n_wide, n_other, n_alu, n_xu, n_f2f=0.0, n_fp64=0.0):
    """Serial-issue model measured by tools/pipe_probe.py: an IMAD.WIDE holds the sub-partition's issue port for
    1/R_wide (4 cycles), every other instruction for one issue slot (1/R_ffma); XU, ALU and FP64 work overlaps."""
    b = {"issue": 1.0 / (n_wide / rates["imad_wide"] + n_other / rates["ffma"]),
         "xu": 1.0 / (n_xu / rates["xu"] + n_f2f / rates["f2f_f32_f64"]) if (n_xu or n_f2f) else float("inf"),
         "alu": rates["lop3"] / n_alu if n_alu else float("inf")}
    if n_fp64:
        b["fp64"] = rates["dadd"] / n_fp64
    return b


# ALGORITHMIC instructions per path-step of the fused kernels, per pipe.  One Philox4x32-10 call =
# 17 IMAD.WIDE (first round's multiplies loop-invariant / warp-uniform) + 20 LOP3; a Box-Muller word = 2 ALU (field
# extraction) + 4 MUFU + FP32 glue.
ALGO = {
    # 8 normals per call; per word (two normals, of which only the SUM enters a constant-variance terminal value:
    # rad (cos a + sin a) = sqrt(2) rad sin(a + pi/4)) 3 FP32 (2 - f, angle FMA, range FMUL) + 1 to accumulate and 3 MUFU
    # (lg2, sqrt, sin)
    "gbm_f32_greeks": {"imad_wide": 17 / 8, "alu": (20 + 4 * 2) / 8, "fp32": 4 * 4 / 8, "xu": 1.5, "loop": 4 / 8},
    # fp64 path state: the same draws, each widened (F2F, XU-rate pipe) and added in fp64 (DADD)
    "gbm_f64_greeks": {"imad_wide": 17 / 8, "alu": (20 + 4 * 2) / 8, "fp32": 4 * 3 / 8 + 1.0, "xu": 2.0, "loop": 4 / 8,
                       "f2f": 1.0, "fp64": 1.0},
    # one pair (Z1, Z2) per step, 4 steps per call; 5 FP32 + 3 MUFU (lg2, sin, cos) to prepare the step's draws -- the
    # Box-Muller radius is never formed: per state sqrt(v+ L) replaces sqrt(v+) and sqrt(L) --; per state 5 FP32 + 1 FMNMX +
    # 1 MUFU.SQRT
    "heston_f32_antithetic": {"imad_wide": 17 / 4, "alu": (20 + 4 * 2) / 4 + 2, "fp32": 5 + 2 * 5, "xu": 3.0 + 2, "loop": 3 / 4},
    # the same step loop: the jumps (drawn per JUMP, not per step) are summed per path outside it (~1 % of the instructions)
    "svj_f32_antithetic": {"imad_wide": 17 / 4, "alu": (20 + 4 * 2) / 4 + 2, "fp32": 5 + 2 * 5, "xu": 3.0 + 2, "loop": 3 / 4},
}
SASS_NAME = {"gbm_f32_greeks": "k_europeanILi0ELb0ELb1EfLb1E", "gbm_f64_greeks": "k_europeanILi0ELb0ELb1EdLb1E",
             "heston_f32_antithetic": "k_europeanILi2ELb1ELb0EfLb1E", "svj_f32_antithetic": "k_europeanILi3ELb1ELb0EfLb1E"}
KERNEL_LABEL = {"gbm_f32_greeks": "k_european<GBM, fp32, greeks>", "gbm_f64_greeks": "k_european<GBM, fp64, greeks>",
                "heston_f32_antithetic": "k_european<HESTON, fp32, antithetic>", "svj_f32_antithetic": "k_european<SVJ, fp32, antithetic>"}


def load_traffic():
    """DRAM bytes per launch of the named kernels, from the ncu --set full captures committed under profiles/ (written by
    tools/ncu_traffic.py from the .ncu-rep of THIS round's kernels); absent -> None (never a stale literal)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    except Exception:  # noqa: BLE001
        return {}


def instruction_roofline(which, rates, achieved, steps_per_call):
    """Instruction roofline of a fused kernel (it moves no data): algorithmic per-path-step instruction counts by
    pipe over the issue rates measured now on this device; the SASS mix of the shipped kernel's hot loop beside it."""
    a = ALGO[which]
    other = a["alu"] + a["fp32"] + a["xu"] + a["loop"] + a.get("f2f", 0.0) + a.get("fp64", 0.0)
    b_algo = pipe_bounds(rates, a["imad_wide"], other, a["alu"], a["xu"], a.get("f2f", 0.0), a.get("fp64", 0.0))
    binding = min(b_algo, key=b_algo.get)
    peak = b_algo[binding]
    tr = load_traffic().get(which, {})
    out = {"bound": binding, "achieved": achieved, "peak": peak, "unit": UNIT, "frac": achieved / peak,
           "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
           "kernel": KERNEL_LABEL[which],
           "kind": "instruction roofline (the kernel moves no data): algorithmic instructions per path-step over issue rates "
                   "measured in this run; the binding pipe is the smallest of pipe_bounds_algorithmic",
           "algorithmic_per_path_step": a, "pipe_bounds_algorithmic": b_algo,
           "peak_source": "tools/probe/libb200mc_probe.so on this device in this run (MEASURED_PEAKS.json holds no pipe rates)"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_mix
        mix = list(sass_mix.mix(SASS_NAME[which]).values())[0]
        if not mix.get("xu"):
            raise RuntimeError("empty mix")
        calls = max(1, round(mix["imad_wide"] / 18.0))
        per = {k: mix.get(k, 0) / (steps_per_call * calls) for k in ("heavy", "alu", "fp32", "xu", "fp64", "uni", "lsu", "ctl", "total", "imad_wide")}
        b_sass = pipe_bounds(rates, per["imad_wide"], per["total"] - per["imad_wide"], per["alu"], per["xu"], 0.0, per["fp64"])
        out.update({"sass_per_path_step": per, "pipe_bounds_sass_mix": b_sass,
                    "frac_of_sass_mix_bound": achieved / min(b_sass.values()),
                    "mix_source": "cuobjdump -sass of libb200mc.so, hot loop (tools/sass_mix.py)"})
    except Exception as e:  # noqa: BLE001
        out["mix_source"] = f"SASS mix unavailable ({type(e).__name__}: {e})"
    return out


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, read+write)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def hbm_roofline(h, torch, rates, n_paths=4_000_000, reps=3):
    """BASELINE cfg4: the path-store kernel (4M paths x 250 steps; 4 or 8 bytes per path-step, written once, output
    far larger than L2) vs the measured copy bandwidth."""
    peak, src = hbm_peak()
    from monte_carlo_option_simulator_b200 import _lib
    out = {}
    tr = load_traffic()
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
            ms = h.timer_end()                      # records the end event and synchronises on it
            if r > 0:
                best = ms if best is None else min(best, ms)
        nbytes = n_paths * (N_STEPS + 1) * esz          # algorithmic bytes: the matrix itself, no padding
        gbs = nbytes / (best * 1e-3) / 1e9
        t = tr.get("paths_" + name, {})
        out[name] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": t.get("dram_bytes_per_launch"), "traffic_source": t.get("source"),
                     "kernel": "k_paths_tma<GBM>" if ld == N_STEPS + 1 else "k_paths_det<GBM>", "path_steps_per_s": n_paths * N_STEPS / (best * 1e-3),
                     "bytes_per_launch": nbytes, "ms": best, "peak_source": src, "shape": [n_paths, N_STEPS + 1], "ld": ld,
                     "note": note}
        del buf
    # the fp32 store carries ~24 instructions per 4 stored bytes (Philox + Box-Muller + the polynomial growth factor + the
    # shared-memory fill): its binding roof is the SAME instruction-issue model as the fused kernel, not HBM -- state both.
    # Algorithmic count per stored value: the draw (2.3 wide, 3.5 ALU, 2.5 FP32, 2 MUFU) + delta FMA + degree-4 expm1 (3 FMA +
    # FMUL) + running product FMA + scaling FMUL + STS (+ the chunk totals) + loop.
    a = {"imad_wide": 18.5 / 8, "alu": (20 + 4 * 2) / 8, "fp32": 4 * 5 / 8 + 7.0, "xu": 2.0, "lsu": 1.1, "loop": 0.6}
    other = a["alu"] + a["fp32"] + a["xu"] + a["lsu"] + a["loop"]
    b = pipe_bounds(rates, a["imad_wide"], other, a["alu"], a["xu"])
    issue_gbs = min(b.values()) * 4.016 / 1e9
    ncu_instr_per_value = 23.3          # smsp__inst_executed.sum * 32 / 1e9 values, profiles/r02_ncu_k_paths_tma_poly4_minb4.txt
    sass_gbs = 4.016 / 1e9 / (a["imad_wide"] / rates["imad_wide"] + (ncu_instr_per_value - a["imad_wide"]) / rates["ffma"])
    out["f32"]["instruction_bound"] = {"algorithmic_per_path_step": a, "pipe_bounds_path_steps_per_s": b,
                                       "as_GB_per_s": issue_gbs, "frac_of_binding_roof": out["f32"]["achieved"] / min(issue_gbs, peak),
                                       "executed_instructions_per_value_ncu": ncu_instr_per_value,
                                       "executed_mix_bound_GB_per_s": sass_gbs, "frac_of_executed_mix_bound": out["f32"]["achieved"] / sass_gbs,
                                       "note": "min(HBM, issue, XU): ~24 issue slots (+3 per IMAD.WIDE) per stored float put the fp32 store "
                                               "on the instruction-issue roof below the HBM roof (ncu: issue port 91 % occupied); the "
                                               "float64-output variant is HBM-bound"}
    return out


# ------------------------------------------------------------------------------------------------ own arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams, _lib
    from monte_carlo_option_simulator_b200.dist import TorchComm, shard_range

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

    def exchange(buf):
        if use_peer:
            h.peer_allreduce(buf.data_ptr(), buf.numel())
        elif world > 1:
            dist.all_reduce(buf)

    def step(i, flags=_lib.GREEKS, npaths=n, off=rank * n, dst=out):
        h.price_european(p, SPOT, T, N_STEPS, npaths, 42 + i, [STRIKE], True, flags, bumps if flags & _lib.GREEKS else None,
                         path_offset=off, out_dev=dst.data_ptr())
        exchange(dst)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_ms(fn, reps=3):
        """best of `reps`: CUDA-event time of fn() on the launching stream, max over ranks"""
        best = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(reps):
            barrier()
            e0.record(stream)
            fn(r)
            e1.record(stream)
            barrier()
            ms_ = max_over_ranks(e0.elapsed_time(e1))
            best = ms_ if best is None else min(best, ms_)
        return best

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
    # When the timed region is shorter than a few nvidia-smi periods the sampler keeps running over ~1 s of the SAME
    # steps issued back to back right after it (not timed): the reported clocks, power and throttle reasons are those
    # of this workload under sustained load.
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
    ms = max_over_ranks(ms)
    value = world * n * N_STEPS * args.steps / (ms * 1e-3)
    sums = out.cpu().numpy()

    # ---- the exchange, checked where it can fail: the same per-rank vector reduced by k_peer_allreduce and by NCCL ------
    allreduce_check = None
    if world > 1:
        local_buf = torch.zeros(_lib.NSUMS, dtype=torch.float64, device="cuda")
        h.price_european(p, SPOT, T, N_STEPS, n, 4242, [STRIKE], True, _lib.GREEKS, bumps, path_offset=rank * n,
                         out_dev=local_buf.data_ptr())
        torch.cuda.synchronize()
        mine = local_buf.clone()
        via_nccl = local_buf.clone()
        dist.all_reduce(via_nccl)
        exchange(local_buf)
        torch.cuda.synchronize()
        a_, b_, m_ = local_buf.cpu().numpy(), via_nccl.cpu().numpy(), mine.cpu().numpy()
        rel = float(np.max(np.abs(a_ - b_) / np.maximum(np.abs(b_), 1e-300)))
        allreduce_check = {"n_total": float(a_[0]), "expected_n_total": float(world * n), "rank0_n": float(m_[0]),
                           "sum_a_exchange": float(a_[1]), "sum_a_nccl": float(b_[1]), "rank0_sum_a": float(m_[1]),
                           "max_rel_diff_vs_nccl_over_17_sums": rel,
                           "ok": bool(a_[0] == world * n and rel < 1e-12 and a_[1] > m_[1] * (world - 0.5))}
        ok_t = torch.tensor([1.0 if allreduce_check["ok"] else 0.0], device="cuda")
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        allreduce_check["ok_on_every_rank"] = bool(ok_t.item() == 1.0)

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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * N_STEPS * e2e_steps / e2e_s

    line = {}
    if not args.no_extras:
        # ---- BASELINE cfg2, fp64 leg: the same step with the path state and payoffs in fp64 ----------------------------
        f64_steps = max(3, min(args.steps, 10))
        step(0, _lib.GREEKS | _lib.FP64)
        ms64 = device_ms(lambda r: [step(100 + r * f64_steps + i, _lib.GREEKS | _lib.FP64) for i in range(f64_steps)], reps=2)
        fp64_value = world * n * N_STEPS * f64_steps / (ms64 * 1e-3)
        s64 = out.cpu().numpy()
        line["cfg2_fp64_greeks"] = {"value": fp64_value, "unit": UNIT, "ms_per_step": ms64 / f64_steps, "steps": f64_steps,
                                    "dtype": "f64", "paths_per_gpu": n,
                                    "check_price": float(np.exp(-p.r * T) * s64[1] / s64[0])}

        # ---- BASELINE cfg5: STRONG scaling -- one pricing call of n_total paths split over the ranks by shard_range ----
        cfg5 = {}
        for n_total in (100_000_000, 1_000_000_000, 10_000_000_000):
            big = n_total > 1_000_000_000                  # 1e10 paths: 1.4 s per call on one GPU -- one timed repetition
            lo, hi = shard_range(n_total, rank, world)
            if not big:
                step(0, 0, hi - lo, lo)
            ms_n = device_ms(lambda r: step(7 + r, 0, hi - lo, lo), reps=1 if big else 3)
            sums5 = out.cpu().numpy()
            entry = {"n_total": n_total, "ms": ms_n, "path_steps_per_s": n_total * N_STEPS / (ms_n * 1e-3),
                     "paths_this_rank": hi - lo, "n_accumulated": float(sums5[0]),
                     "price": float(np.exp(-p.r * T) * sums5[1] / sums5[0])}
            if world > 1:
                # the same call on ONE GPU, no exchange, timed in this run on every rank (max): the strong-scaling reference
                loc = torch.zeros(_lib.NSUMS, dtype=torch.float64, device="cuda")

                def single(r, loc=loc, n_total=n_total):
                    h.price_european(p, SPOT, T, N_STEPS, n_total, 7 + r, [STRIKE], True, 0, None, path_offset=0, out_dev=loc.data_ptr())
                if not big:
                    single(0)
                ms_1 = device_ms(single, reps=1 if big else 2)
                entry.update({"single_gpu_ms_same_run": ms_1, "efficiency": ms_1 / (world * ms_n),
                              "overhead_ms_vs_ideal": ms_n - ms_1 / world})
            eng5 = MonteCarloEngine(p, n_total, N_STEPS, 42, use_sobol=False, use_antithetic=False, use_control_variate=False,
                                    rng="philox", handle=h, comm=comm)
            if not big:
                eng5.price(SPOT, STRIKE, T, True)
            barrier()
            t0 = time.perf_counter()
            r5 = eng5.price(SPOT, STRIKE, T, True)
            barrier()
            entry["api_price_ms"] = max_over_ranks(time.perf_counter() - t0) * 1e3
            entry["api_price"] = r5["price"]
            cfg5[f"{n_total:.0e}".replace("+0", "").replace("+", "")] = entry
        line["cfg5_strong_scaling"] = {"what": "one European call priced with n_total paths x 250 steps, global path range split "
                                               "contiguously over the ranks, one exchange of 17 sums; CUDA events, max over ranks, best of 3 (1e10: one repetition)",
                                       "runs": cfg5}

        # ---- BASELINE cfg4: 4M x 251 path store SHARDED over the ranks (no exchange) + VaR / CVaR over the shards -------
        n4 = 4_000_000
        lo4, hi4 = shard_range(n4, rank, world)
        rows = hi4 - lo4
        peak, peak_src = hbm_peak()
        cfg4 = {"n_paths_total": n4, "rows_this_rank": rows, "peak_GB_per_s": peak, "peak_source": peak_src}
        for name, dt_np, esz in (("f32", np.float32, 4), ("f64", np.float64, 8)):
            buf = torch.empty(rows * (N_STEPS + 1) * esz, dtype=torch.uint8, device="cuda")
            best = None
            for r in range(4):
                barrier()
                h.timer_begin()
                h.generate_paths(p, SPOT, T, N_STEPS, rows, 42, 0, dt_np, lo4, N_STEPS + 1, out_dev=buf.data_ptr())
                ms_ = h.timer_end()                 # end event recorded on the launching stream, then synchronised
                if r:
                    best = ms_ if best is None else min(best, ms_)
            mine_gbs = rows * (N_STEPS + 1) * esz / (best * 1e-3) / 1e9
            per_rank = [mine_gbs]
            if world > 1:
                gl = [None] * world
                dist.all_gather_object(gl, mine_gbs)
                per_rank = [float(x) for x in gl]
            slowest = max_over_ranks(best)
            cfg4[name] = {"per_rank_GB_per_s": per_rank, "min_rank_frac_of_peak": min(per_rank) / peak,
                          "aggregate_GB_per_s": n4 * (N_STEPS + 1) * esz / (slowest * 1e-3) / 1e9, "ms_slowest_rank": slowest,
                          "path_steps_per_s": n4 * N_STEPS / (slowest * 1e-3)}
            if name == "f32":
                # paths -> discounted option P&L from the last column (device) -> tail metrics over the sharded vector
                pnl = torch.empty(max(rows, 1), dtype=torch.float64, device="cuda")
                from monte_carlo_option_simulator_b200.risk import compute_risk_metrics, compute_risk_metrics_sharded

                def tail():
                    h.option_pnl(buf.data_ptr() + N_STEPS * 4, rows, STRIKE, True, float(np.exp(-p.r * T)), 374.0712289657911,
                                 pnl.data_ptr(), dtype_in=np.float32, stride=N_STEPS + 1)
                    if world > 1:
                        return compute_risk_metrics_sharded((pnl.data_ptr(), rows, np.float64), 0.99, comm=comm, handle=h)
                    return dict(zip(("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std"),
                                    h.risk_metrics(pnl.data_ptr(), 0.99, n=rows, dtype=np.float64)))
                tail()
                barrier()
                t0 = time.perf_counter()
                met = tail()
                torch.cuda.synchronize()
                cfg4["pnl_and_tail_metrics_ms"] = max_over_ranks(time.perf_counter() - t0) * 1e3
                cfg4["var_99"], cfg4["cvar_99"] = met["var"], met["cvar"]
                del pnl
            del buf
        cfg4["note"] = ("every rank stores rows [lo, hi) of the GLOBAL path index in its own HBM (same rows as a 1-GPU run), timer = CUDA "
                        "events on the launching stream with a synchronise before the read; VaR/CVaR must not depend on the GPU count")
        line["cfg4_sharded_path_store"] = cfg4

    # BASELINE cfg3, benchmark reading: 64 strikes x 16 expiries, 1M paths PER CELL (disjoint counter ranges), cells dealt
    # round-robin to the ranks, no path-level collective (one all-reduce gathers the 1024 sum vectors)
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
        dt3 = max_over_ranks(time.perf_counter() - t0)
        work3 = 64 * 1_000_000 * float(g3["num_steps"].sum())
        cfg3 = {"seconds": dt3, "cells": 1024, "path_steps": work3, "path_steps_per_s": work3 / dt3,
                "reading": "independent cells: 1M paths per (expiry, strike) cell, cells sharded over the ranks",
                "atm_1y_price": float(g3["prices"][7, 32]), "atm_1y_std_error": float(g3["std_errors"][7, 32])}

    if rank == 0:
        disc = float(np.exp(-p.r * T))
        head = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
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
                          "n_total": float(sums[0]), "expected_n_total": float(world * n),
                          "api_delta": last[0]["pathwise"], "api_gamma": last[2]["gamma"], "allreduce": allreduce_check}}
        head.update(line)
        line = head
        if world == 1 and not args.no_extras:
            rates = measured_rates(h)
            line["roofline"] = instruction_roofline("gbm_f32_greeks", rates, value, 8.0)
            line["roofline"]["measured_rates_ops_per_s"] = rates
            line["roofline"]["rng_only_path_steps_per_s"] = 8.0 * rates["philox_bm_calls"]
            line["roofline_fp64"] = instruction_roofline("gbm_f64_greeks", rates, line["cfg2_fp64_greeks"]["value"], 8.0)
            try:
                line["roofline_hbm"] = hbm_roofline(h, torch, rates)
            except Exception as e:  # noqa: BLE001
                line["roofline_hbm"] = {"error": str(e)}
            extras = {}

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
            extras["fp32_heston_price_only"] = rate(SVJParams(lambda_j=0.0), 22500.0, n // 4, 0, [22500.0])
            extras["fp32_svj_price_only"] = rate(SVJParams(), 22500.0, n // 4, 0, [22500.0])
            # SURVEY 8(d) counts the antithetic twin as a path of its own; this file counts a PAIR as one path everywhere
            # else, so the pair rates are also given with the twin counted (2 x)
            for k in ("fp32_antithetic", "fp32_heston_antithetic", "fp32_svj_antithetic"):
                extras[k + "_twin_counted_as_a_path"] = 2.0 * extras[k]
            # the reference's default model (SVJParams(): Heston + jumps) and its jump-free special case: antithetic pairs,
            # a pair counted as ONE path in the rooflines
            line["roofline_heston"] = instruction_roofline("heston_f32_antithetic", rates, extras["fp32_heston_antithetic"], 4.0)
            line["roofline_svj"] = instruction_roofline("svj_f32_antithetic", rates, extras["fp32_svj_antithetic"], 4.0)
            # BASELINE cfg3, API-parity reading: 64 strikes x 16 expiries, 1M paths per expiry shared across strikes
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
                gq = np.random.default_rng(1)
                kk, tt = 22500.0 * gq.uniform(0.8, 1.2, 2048), gq.uniform(0.05, 1.0, 2048)
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
                x4 = torch.randn(4_000_000, dtype=torch.float64, device="cuda")
                h.risk_metrics(x4.data_ptr(), 0.99, n=4_000_000, dtype=np.float64)
                l_r = h.launches
                t0 = time.perf_counter()
                h.risk_metrics(x4.data_ptr(), 0.99, n=4_000_000, dtype=np.float64)
                nxt["tail_metrics_4M_values"] = {"seconds": time.perf_counter() - t0, "launches": int(h.launches - l_r)}
                line["next_rows"] = nxt
            except Exception as e:  # noqa: BLE001
                line["next_rows"] = {"error": str(e)}
            try:
                arm = make_cpu_arm(50_000)
                arm.step()                                          # JIT / first-touch warm-up, not timed
                dtc = _timed(arm.step)
                line["cpu_baseline"] = cpu_baseline_block(arm, arm.n * N_STEPS / dtc, dtc, 1, arm.readings(3))
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()


_JSON_FD = None


def emit(line: dict):
    """The ONE line of stdout.  main() points file descriptor 1 at stderr for the rest of the run, so that nothing a
    library prints from C (NCCL's version banner on its first collective, Numba / OpenMP notices) can land in front of it."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps (100 x 1.4 ms: the clock samples fall inside the region)")
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
