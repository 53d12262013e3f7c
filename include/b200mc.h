/*
 * b200mc.h -- C ABI of libb200mc.so, the B200 (sm_100a) Monte Carlo pricing core.
 *
 * This is the drop-in boundary for ONE hot path of Jay14090/Monte-Carlo-Option-Simulator: the SVJ/GBM
 * path simulation and the payoff / Greek / VaR reductions built on it.  The reference has no FFI of its
 * own (pure Python + Numba); each entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  The Python host mirror that binds these with ctypes is
 * monte_carlo_option_simulator_b200/_lib.py; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - Plain C types only.  Every buffer is allocated by the caller; the library never returns memory it owns.
 *   - Every function returns 0 on success or a B200MC_E* code; it never aborts or throws across the ABI.
 *     b200mc_last_error(h) gives the message of the last failure on that handle (NULL handle: the message of
 *     the last failed b200mc_create on this thread).  The string is valid until the next call on the handle.
 *   - A handle is bound to one CUDA device and owns one stream plus scratch; it is NOT thread-safe (one call
 *     in flight per handle), matching the reference where every caller is serial (engine/app.py:130-236,
 *     engine/calibration.py:202).  Callers that share a handle between threads serialise their calls (the Python
 *     binding does: a lock per handle); several handles on one device may be used concurrently, one per thread.
 *     Calls are synchronous unless the name ends in _async.
 *   - Host pointers may be pageable or pinned.  *_dev variants take device pointers valid on the handle's
 *     device and run on the handle's stream.
 *   - There is no CPU fallback: without a usable sm_100 device b200mc_create fails with B200MC_ENODEVICE.
 *
 * Random numbers (fused modes): counter-based Philox4x32-10, key = (seed_lo, seed_hi),
 * counter = (path_lo, path_hi, block, stream) with `path` the GLOBAL path index (path_offset + i), so a
 * path draws the same numbers whatever the launch geometry or the number of GPUs.  Every 32-bit output word w
 * yields one Box-Muller pair BM(w) = (R cos a, R sin a):  u1 = 2 - f(w >> 9) in (0,1],  R = sqrt(-2 ln u1),
 * a = 2 pi f((w * 0x9E3779B9 mod 2^32) >> 9) - 3 pi,  f(m) = the float in [1,2) with mantissa m  (fp32, MUFU
 * lg2/sqrt/sin/cos; the angle field is the word hashed by the golden-ratio multiplier, so that (radius, angle) is a
 * well-spread rank-1 lattice instead of two overlapping bit fields).  Streams:
 *   B200MC_STREAM_GBM    block j -> steps 8j..8j+7: word i gives the normals of steps 8j+2i and 8j+2i+1
 *   B200MC_STREAM_HESTON block j -> steps 4j..4j+3: word i gives (Z1, Z2) of step 4j+i
 *   B200MC_STREAM_SVJ    block j -> steps 4j..4j+3 like B200MC_STREAM_HESTON (the diffusion of the SVJ model)
 *   B200MC_STREAM_SVJ_JUMP  the jump times and sizes of the SVJ model.  The reference tests U < p = lambda_j dt at every
 *                        step (engine/monte_carlo.py:233); the same process is drawn here as geometric GAPS, one uniform
 *                        per jump: block k -> jumps 2k and 2k+1 of the path, (w0 -> gap, w1 -> size), (w2 -> gap,
 *                        w3 -> size); gap = floor(lg2(U) / lg2(1 - p)) jump-free steps before the jump,
 *                        U = ((float)w + 0.5) 2^-32 (fp32); size Z_jump_size = first member of BM(w).
 *                        b200mc_dump_normals(B200MC_ZJUMP_U) turns the jump times back into a per-step uniform array
 *                        (p U' at the jump steps, p + (1 - p) U'' elsewhere) that makes the reference fire the same jumps.
 * b200mc_dump_normals returns exactly the values the fused kernels use, so the reference (or the oracle) can be
 * fed identical draws.
 */
#ifndef B200MC_H
#define B200MC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MC_VERSION 100 /* 0.1.0 */

/* status codes */
#define B200MC_OK        0
#define B200MC_EINVAL    1 /* bad argument (NULL pointer, n <= 0, ...)           */
#define B200MC_ENODEVICE 2 /* no CUDA device / not an sm_100 part                */
#define B200MC_ECUDA     3 /* CUDA runtime error (message has the CUDA string)   */
#define B200MC_ENOMEM    4 /* device or pinned-host allocation failed            */

/* flags for the fused entry points */
#define B200MC_ANTITHETIC 0x1u  /* also evolve the twin (-Z1,-Z2,U,-Zjs): engine/monte_carlo.py:318-324 */
#define B200MC_GREEKS     0x2u  /* fill the bump / pathwise accumulators of b200mc_sums                  */
#define B200MC_FP64       0x4u  /* evolve the path state and the payoff in fp64 (draws are unchanged)    */
#define B200MC_FORCE_SVJ  0x8u  /* use the general SVJ kernel even when xi == 0 and lambda_j == 0        */
#define B200MC_WIDE_RNG   0x10u /* validation twin of the generator (b200mc_price_european, constant variance, one
                                 * strike, no GREEKS): one Box-Muller pair per TWO words (radius and angle share no bit),
                                 * block j -> steps 4j..4j+3.  Half the normals per Philox call; never the default.      */

#define B200MC_STREAM_GBM    0u
#define B200MC_STREAM_HESTON 1u
#define B200MC_STREAM_SVJ    2u
#define B200MC_STREAM_HEDGE  3u   /* b200mc_hedge_walk: GBM layout, counter path = scenario index */
#define B200MC_STREAM_SVJ_JUMP 4u /* jump gaps and sizes of the SVJ model (not a b200mc_dump_normals selector) */
#define B200MC_STREAM_FILL   5u   /* b200mc_dump_normals only: the uniforms of the jump-free steps of ZJUMP_U */

/* which array b200mc_dump_normals returns */
#define B200MC_Z1         0
#define B200MC_Z2         1
#define B200MC_ZJUMP_U    2
#define B200MC_ZJUMP_SIZE 3

#define B200MC_F32 0
#define B200MC_F64 1

typedef struct b200mc_handle b200mc_handle;

/* Field set of SVJParams, engine/models.py:31-44. */
typedef struct {
    double v0, r, q, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j;
} b200mc_svj_params;

/* Bump sizes for the CRN finite-difference Greeks (all evaluated on the SAME draws as the base path). */
typedef struct {
    double spot_bump; /* relative: spot*(1 +- b); engine/greeks.py:54,79-80,168,179 (0.01)  */
    double v0_up;     /* absolute bumped v0; engine/greeks.py:124 (v0 + 0.01)               */
    double v0_dn;     /* engine/greeks.py:125 (max(v0 - 0.01, 0.001))                       */
    double r_up;      /* absolute bumped rate; engine/greeks.py:235 (r + 1e-4)              */
    double r_dn;      /* engine/greeks.py:239 (max(r - 1e-4, 0))                            */
} b200mc_bumps;

/*
 * Per-strike sums over the paths of one launch (all fp64, all UNDISCOUNTED payoffs).
 * a_i = payoff of the primary path, b_i = payoff of its antithetic twin (0 when ANTITHETIC is off).
 * Everything MonteCarloEngine.price / price_batch return derives from n, sum_a, sum_b, sum_aa, sum_bb,
 * sum_ab (engine/monte_carlo.py:327-373, :416-448); the Greek fields feed GreeksEngine (engine/greeks.py).
 * Bumped payoffs use the primary path only (the reference's Greeks never use antithetic draws).
 */
typedef struct {
    double n;            /* number of primary paths accumulated                                      */
    double sum_a, sum_b, sum_aa, sum_bb, sum_ab;
    double sum_s;        /* sum of S_T (pair average with ANTITHETIC): true control variate, E known */
    double sum_ss;       /* sum of that quantity squared                                             */
    double sum_ps;       /* sum of (pair-averaged payoff) * (pair-averaged S_T)                      */
    /* --- filled only with B200MC_GREEKS (else 0) ------------------------------------------------- */
    double sum_pw_delta; /* sum 1{ITM} S_T / S0                  engine/greeks.py:71-76              */
    double sum_spot_up;  /* sum payoff(S_T (1+b))                engine/greeks.py:79,83 / :184,189   */
    double sum_spot_dn;  /* sum payoff(S_T (1-b))                engine/greeks.py:80,84 / :185,190   */
    double sum_v0_up;    /* sum payoff(S_T | v0 = v0_up)         engine/greeks.py:136-141,150        */
    double sum_v0_dn;    /* sum payoff(S_T | v0 = v0_dn)         engine/greeks.py:142-147,151        */
    double sum_r_up;     /* sum payoff(S_T e^{(r_up-r)T})  (discount with r_up on the host)          */
    double sum_r_dn;     /* sum payoff(S_T e^{(r_dn-r)T})                                            */
    double sum_pw_vega;  /* sum 1{ITM} S_T dlogS_T/dsigma, constant-variance (GBM) case only; new    */
} b200mc_sums;

/* ---- lifetime --------------------------------------------------------------------------------------- */
int b200mc_version(void);
int b200mc_create(int device, b200mc_handle **out);
int b200mc_destroy(b200mc_handle *h);
const char *b200mc_last_error(const b200mc_handle *h);
/* sm_count, max SM clock (kHz), HBM bytes, compute capability (major*10+minor) of the handle's device */
int b200mc_device_info(const b200mc_handle *h, int *sm_count, int *sm_clock_khz, uint64_t *hbm_bytes, int *cc);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t b200mc_launch_count(const b200mc_handle *h);
/* cudaStream_t of the handle as an integer, so a caller can record CUDA events on the launching stream */
uint64_t b200mc_stream(const b200mc_handle *h);
/* Adopt a caller-owned cudaStream_t (e.g. torch's current stream, so that a following NCCL all-reduce is ordered
 * after the kernels).  The handle's own stream is drained and destroyed; the caller keeps ownership of the new one. */
int b200mc_set_stream(b200mc_handle *h, uint64_t stream);
int b200mc_synchronize(b200mc_handle *h);

/* ---- the exchange step of the multi-GPU path over NVLink peer memory ---------------------------------------
 * One process per GPU.  b200mc_peer_create allocates this rank's exchange buffer and returns its 64-byte CUDA IPC handle;
 * the caller all-gathers the handles of all ranks (any transport: torch.distributed, MPI, a file) and passes them, in rank
 * order, to b200mc_peer_connect, which maps the peers' buffers.  b200mc_peer_allreduce then sums data_dev[n_doubles]
 * (DEVICE memory of this rank, e.g. the b200mc_sums a b200mc_price_european_async call just wrote) over all ranks IN
 * PLACE: one single-CTA kernel on the handle's stream that stores the vector into every rank's buffer, raises a flag,
 * waits for the flags of all ranks and adds the vectors in rank order (bitwise the same result on every rank).  A
 * collective: all ranks call it, in the same order.  A peer that never arrives turns the result into NaN after 20 s. */
#define B200MC_PEER_MAX_RANKS   16
#define B200MC_PEER_MAX_DOUBLES 4352            /* 256 strikes x 17 sums */
int b200mc_peer_create(b200mc_handle *h, unsigned char ipc_handle_out[64]);
int b200mc_peer_connect(b200mc_handle *h, int rank, int world, const unsigned char *all_handles /* [world][64] */);
int b200mc_peer_allreduce(b200mc_handle *h, double *data_dev, int32_t n_doubles);
int b200mc_peer_close(b200mc_handle *h);
/* compute_risk_metrics (engine/risk.py:117-155) over a vector SHARDED across the ranks of the peer connection: every rank
 * passes its own shard (n_local may be 0) and receives the GLOBAL metrics (out as b200mc_risk_metrics).  The shards never
 * move: the radix select runs on every rank's keys and only 3 doubles, 8 x 256-bin histograms and 6 doubles are all-reduced,
 * on the device, between the kernels.  A collective: all ranks call it. */
int b200mc_risk_metrics_sharded(b200mc_handle *h, const void *pnl_local, int64_t n_local, int dtype, int on_device,
                                double confidence, double out[8]);

/* ---- a1: deterministic "given normals" mode ------------------------------------------------------------
 * Drop-in for _simulate_svj_paths_numba(S0, v0, r, q, T, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j,
 * Z1, Z2, Z_jump, Z_jump_size, num_steps, record_paths), engine/monte_carlo.py:189-243.
 * Z* are C-contiguous float64 [n_paths, n_steps] HOST arrays; S_final, v_final float64 [n_paths];
 * all_paths float64 [n_paths, n_steps + 1] (column 0 = S0) or NULL when record_paths == 0.  fp64 arithmetic,
 * the reference's operation order (multiplicative S *= exp(..)); expect ~1e-13 relative agreement.
 * Arrays the parameters make irrelevant are not read (nor copied to the device): Z2 when xi == 0, and Z_jump /
 * Z_jump_size when lambda_j * dt <= 0 (Z_jump is a uniform in [0, 1), so the test at :233 cannot fire). */
int b200mc_simulate_given_normals(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                  int64_t n_paths, int32_t n_steps,
                                  const double *Z1, const double *Z2,
                                  const double *Z_jump, const double *Z_jump_size,
                                  int record_paths, double *S_final, double *v_final, double *all_paths);
/* record_paths is a flag word: B200MC_GIVEN_RECORD (= 1, the reference's record_paths=True) and / or B200MC_GIVEN_NEGATE:
 * evaluate the antithetic twin, i.e. use -Z1, -Z2, Z_jump, -Z_jump_size (engine/monte_carlo.py:318-324) WITHOUT the caller
 * materialising negated copies of the arrays (the negation is exact). */
#define B200MC_GIVEN_RECORD 1
#define B200MC_GIVEN_NEGATE 2
/* Same with DEVICE pointers (inputs already resident in HBM); asynchronous on the handle's stream. */
int b200mc_simulate_given_normals_dev(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                      int64_t n_paths, int32_t n_steps,
                                      const double *Z1, const double *Z2,
                                      const double *Z_jump, const double *Z_jump_size,
                                      int record_paths, double *S_final, double *v_final, double *all_paths);

/* ---- a2: NumPy's pseudo-random front end on the device, bit for bit --------------------------------------
 * Replaces np.random.default_rng(seed).standard_normal((n, steps)) x 3 and default_rng(seed + 1).random((n, steps))
 * (engine/monte_carlo.py:301-308, engine/greeks.py:33-41, :458-462) for rng="reference" runs: PCG64 + NumPy's 256-layer
 * Ziggurat (numpy 2.3.5, random_standard_normal), the data-dependent stream resolved in parallel (csrc/np_normal.cu).
 * state = {state_hi, state_lo, inc_hi, inc_lo} of default_rng(seed).bit_generator (the SeedSequence hash stays in NumPy);
 * first_raw = generator outputs already consumed (consecutive calls on one generator chain through *raws_consumed);
 * out: n doubles, device pointer when on_device else host.  Successive standard_normal calls on one generator are ONE
 * stream, so Z1 | Z2 | Z_jump_size of the reference are a single call with n = 3 * n_paths * n_steps. */
#define B200MC_NUMPY_RANDOM          0   /* Generator.random():          one output per double                    */
#define B200MC_NUMPY_STANDARD_NORMAL 1   /* Generator.standard_normal(): 1.0215 outputs per double on average     */
int b200mc_numpy_fill(b200mc_handle *h, const uint64_t state[4], uint64_t first_raw, int64_t n, int kind, int on_device,
                      double *out, uint64_t *raws_consumed);
/* The Ziggurat tables in use (numpy's ki_double / wi_double / fi_double).  Host only, no device needed. */
int b200mc_numpy_ziggurat_tables(uint64_t ki[256], double wi[256], double fi[256]);

/* ---- a2+a1+a3/a4/a6-a9 fused: Philox draws in registers, payoff and Greek sums reduced on chip ----------
 * Replaces the RNG front end + both kernel runs + the NumPy reductions of MonteCarloEngine.price
 * (engine/monte_carlo.py:273-375), .price_batch (:377-450) and GreeksEngine.delta/vega/gamma
 * (engine/greeks.py:53-203) for the pseudo-random case.  Simulates global paths
 * [path_offset, path_offset + n_paths) for n_steps steps of dt = T / n_steps (the caller applies the
 * reference's steps rule, monte_carlo.py:287) and writes one b200mc_sums per strike into out[n_strikes]
 * (host memory).  `bumps` may be NULL unless B200MC_GREEKS is set.  is_call: 1 call, 0 put.
 * No path matrix touches HBM.  Sums of disjoint path ranges add (multi-GPU: all-reduce the structs).
 * The call returns after the kernel has finished; the sums reach the host without a copy command (the kernel's last
 * CTA stores them into a pinned, device-mapped landing buffer of the handle, which is then copied into `out`): a
 * 10k-path x 50-step call takes ~20 us end to end on B200.  B200MC_RESULT=copy in the environment selects a device
 * buffer + cudaMemcpyAsync instead (the same numbers, ~8 us more per call). */
int b200mc_price_european(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                          int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                          const double *strikes, int32_t n_strikes, int is_call, uint32_t flags,
                          const b200mc_bumps *bumps, b200mc_sums *out);
/* Asynchronous form: launches on the handle's stream, results land in out_dev (DEVICE, n_strikes structs). */
int b200mc_price_european_async(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                                const double *strikes, int32_t n_strikes, int is_call, uint32_t flags,
                                const b200mc_bumps *bumps, b200mc_sums *out_dev);

/* ---- a11 / SURVEY 8(f)-1,2: many independent pricing problems in one launch over a (cell x path) grid ------
 * The reference prices scenario ladders and optimiser candidates as loops of MonteCarloEngine.price():
 * StressTestEngine (engine/risk.py:33-111, 13 calls), the premium of every HedgingBacktest scenario
 * (engine/risk.py:264-273, num_scenarios calls), the calibration objectives (engine/calibration.py:78-89,
 * 119-130).  A cell is one such call: its own parameters, spot, maturity, steps, paths, seed, option type.
 * All cells of a call share n_strikes and flags (ANTITHETIC, FP64, FORCE_SVJ; GREEKS is rejected);
 * strikes is [n_cells][n_strikes] (host), out is [n_cells][n_strikes] b200mc_sums (host, or device memory when
 * on_device != 0: then the call is asynchronous on the handle's stream).  Cell i's sums equal those of
 * b200mc_price_european(cell i's arguments, strikes[i], ...) -- same draws, same per-path arithmetic -- up to the
 * order of the fp64 additions; the Greek fields are 0. */
typedef struct b200mc_cell {
    b200mc_svj_params params;
    double S0, T;
    int64_t n_paths;
    uint64_t seed, path_offset;
    int32_t n_steps;
    int32_t is_call;
} b200mc_cell;
int b200mc_price_cells(b200mc_handle *h, const b200mc_cell *cells, int32_t n_cells, const double *strikes,
                       int32_t n_strikes, uint32_t flags, int on_device, b200mc_sums *out);

/* The daily delta-hedging walk of HedgingBacktest.run_backtest (engine/risk.py:278-316) for n_scenarios scenarios:
 * Black-Scholes delta at sigma = sqrt(v0) (engine/monte_carlo.py:45-55), rebalancing cost |trade| S cost_bps / 10000
 * with cost_bps = txn_cost_bps + slippage_bps, one GBM step per day, settlement against the option payoff; fp64 in the
 * reference's operation order.  premiums: [n_scenarios] cash received at t = 0 (risk.py:271-273; NULL = 0).
 * Z: [n_scenarios][n_days] standard normals (HOST; the reference's default_rng(seed).standard_normal() sequence) or
 * NULL: then the normals are Philox draws, counter = (scenario_offset + i, day / 8, B200MC_STREAM_HEDGE), key = seed,
 * exactly what b200mc_dump_normals(seed, scenario_offset, n, n_days, B200MC_STREAM_HEDGE, B200MC_Z1) returns.
 * Outputs (HOST): final_pnl[n_scenarios]; txn_cost[n_scenarios] total rebalancing cost (may be NULL). */
int b200mc_hedge_walk(b200mc_handle *h, const b200mc_svj_params *p, double S0, double strike, double T, int is_call,
                      int32_t n_days, int64_t n_scenarios, double cost_bps, const double *premiums, const double *Z,
                      uint64_t seed, uint64_t scenario_offset, double *final_pnl, double *txn_cost);

/* Black-Scholes implied volatilities of n options in one launch -- the double loop of
 * extract_iv_surface (engine/surface.py:69-126) and the smile loop of engine/app.py:226-234, whose every iteration is
 * implied_vol(price, S, K, T, r, q, is_call, lo = 0.001, hi = 5.0) = brentq on bs_call_price / bs_put_price
 * (engine/surface.py:22-66, xtol 1e-8).  prices, strikes, maturities: [n] float64; is_call: [n] int32 (HOST).
 * out_iv[n]: the root in [lo, hi] (bracketed Newton, converged to ~1e-13; brentq's answer lies within its xtol of it),
 * NaN where the reference returns None (price(lo) - p and price(hi) - p of the same sign, or NaN inputs).  Where p
 * equals price(lo) or price(hi) to fp64 rounding that sign test hangs on the last bit of the normal CDF (SciPy's ndtr
 * there, CUDA's normcdf here). */
int b200mc_implied_vol(b200mc_handle *h, int64_t n, const double *prices, const double *strikes,
                       const double *maturities, const int32_t *is_call, double S, double r, double q,
                       double lo, double hi, double *out_iv);

/* ---- a working quasi-Monte Carlo front end (new; the reference's own use_sobol path is degenerate and
 * stays reproducible on the host) ---------------------------------------------------------------------
 * Scrambled Sobol points generated on the device, bitwise those of scipy.stats.qmc.Sobol(d, scramble=True, seed): the
 * caller passes that engine's scrambled direction numbers sv[n_dims][bits] and digital shift[n_dims] (uint32, HOST);
 * point n = shift ^ XOR over the set bits b of gray(n) of sv[.][b]; u = clip(x 2^-bits, 1e-10, 1 - 1e-10),
 * z = normcdfinv(u) (engine/monte_carlo.py:80-84).  Dimension layout, s = n_steps: [0, s) Z1 and [s, 2s) Z2 in
 * Brownian-bridge order (dimension 0 = W_T, then interval midpoints breadth first: a correct bridge), [2s, 3s) jump
 * sizes and [3s, 4s) jump uniforms in time order; n_dims must cover the blocks the parameters need (s if xi == 0 and
 * lambda_j == 0, 2s if lambda_j == 0, else 4s).  Paths [path_offset, path_offset + n_paths) of the sequence.
 * b200mc_qmc_normals: the step normals / uniforms of one block, float64 [n_paths][n_steps] on the HOST (which =
 * B200MC_Z1 | Z2 | ZJUMP_U | ZJUMP_SIZE), for checks against SciPy.
 * b200mc_price_european_qmc: those draws, device resident, through the fp64 recurrence of
 * b200mc_simulate_given_normals_dev, terminal values reduced on the device; flags: ANTITHETIC only; out[n_strikes]
 * with the price sums of b200mc_sums (Greek fields 0). */
/* One placement of a Brownian-bridge construction: W[t] = W[l] + ((W[r] - W[l]) * a) / b + sd * z[dim], in construction
 * order; index 0 is the known W_0 = 0, t in [1, n_steps], every t exactly once, l and r placed earlier (or 0). */
typedef struct b200mc_bridge_node {
    int32_t t, l, r, dim;
    double a, b, sd;
} b200mc_bridge_node;
int b200mc_qmc_normals(b200mc_handle *h, int64_t n_paths, uint64_t path_offset, int32_t n_steps, const uint32_t *sv,
                       const uint32_t *shift, int32_t n_dims, int32_t bits, int which, double *out);
int b200mc_price_european_qmc(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                              int64_t n_paths, uint64_t path_offset, const uint32_t *sv, const uint32_t *shift,
                              int32_t n_dims, int32_t bits, const double *strikes, int32_t n_strikes, int is_call,
                              uint32_t flags, b200mc_sums *out);

/* NumPy's default_rng(seed).random(n), bit for bit, on the device: state = {state_hi, state_lo, inc_hi, inc_lo} of the PCG64
 * bit generator (NumPy: default_rng(seed).bit_generator.state["state"]), `first` = how many outputs to skip.  out: float64
 * [n] on the HOST.  (The reference draws its jump uniforms this way, engine/monte_carlo.py:308.) */
int b200mc_pcg64_random(b200mc_handle *h, const uint64_t state[4], uint64_t first, int64_t n, double *out);
/* Terminal spots of paths driven by the Sobol point set through a CALLER-SUPPLIED bridge table (nodes[n_nodes], n_nodes ==
 * n_steps; NULL = the built-in correct bridge): the device-side form of the reference's own use_sobol=True front end
 * (engine/monte_carlo.py:290-299).  With the table of the reference's brownian_bridge_reorder (:88-145,172-183 -- degenerate,
 * SURVEY section 0 quirk 1; monte_carlo.reference_bridge_nodes builds it) and Z_jump_host = default_rng(seed + 1).random((n,
 * steps)) (:308; HOST, [n_paths][n_steps]) -- or, with Z_jump_host NULL, pcg64_state[4] = that generator's state, from which
 * the same uniforms are produced on the device (both may be NULL when lambda_j == 0) -- the outputs are those of the reference's default
 * price() inputs (~1e-12), without its seconds of host work.  Blocks of the point set: Z1, Z2 (bridged), jump sizes (plain).
 * flags: ANTITHETIC -> also S_anti (the -Z1, -Z2, U, -Zjs twin).  S_final / S_anti: float64 [n_paths] on the HOST. */
int b200mc_qmc_terminal(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                        int64_t n_paths, uint64_t path_offset, const uint32_t *sv, const uint32_t *shift, int32_t n_dims,
                        int32_t bits, const b200mc_bridge_node *nodes, int32_t n_nodes, const double *Z_jump_host,
                        const uint64_t *pcg64_state, uint32_t flags, double *S_final, double *S_anti);

/* Terminal values of the fused simulation (deterministic-mode parity of the fused kernels, and the terminal
 * P&L vector for compute_risk_metrics, engine/risk.py:117).  S_T / S_T_anti / v_T are [n_paths] of `dtype`
 * (B200MC_F32 / B200MC_F64); any may be NULL.  on_device != 0: pointers are device memory. */
int b200mc_simulate_terminal(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                             int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                             uint32_t flags, int dtype, int on_device,
                             void *S_T, void *S_T_anti, void *v_T);

/* ---- a5 / path-storing mode ------------------------------------------------------------------------------
 * MonteCarloEngine.get_sample_paths (engine/monte_carlo.py:452-471) at scale: the full path matrix
 * [n_paths, n_steps + 1], row-major with leading dimension ld >= n_steps + 1 (elements), column 0 = S0
 * (:216-217,241).  dtype selects the element type of `out` (B200MC_F32 / B200MC_F64), B200MC_FP64 the precision of
 * the path state.  All variants stage tiles in shared memory.  Deterministic variance (GBM) with ld == n_steps + 1: a
 * CTA owns 32 paths, its warps split the time axis, and the finished 32 x (n_steps + 1) tile -- which has exactly
 * the global layout -- leaves with ONE TMA bulk store; padded rows: 128-bit row stores; Heston / SVJ: per-row rings
 * of aligned 64-byte windows.  Only B200MC_FP64 and B200MC_FORCE_SVJ are accepted in flags. */
int b200mc_generate_paths(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                          int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                          uint32_t flags, int dtype, int on_device, void *out, int64_t ld);

/* ---- a10: tail metrics -----------------------------------------------------------------------------------
 * compute_risk_metrics(returns, confidence), engine/risk.py:117-155 (+ _hill_estimator :158-173), with the
 * same index conventions (cutoff = int(n (1 - c)), var = -sorted[cutoff], cvar = -mean(sorted[:cutoff])).
 * out[8] = { var, cvar, skewness, kurtosis, excess_kurtosis, tail_index, mean, std }.  Order statistics are
 * found by radix select (exact, no sort). */
int b200mc_risk_metrics(b200mc_handle *h, const void *pnl, int64_t n, int dtype, int on_device,
                        double confidence, double out[8]);

/* Discounted option P&L on the device: pnl[i] = discount * payoff(S[i * stride]) - premium, S and pnl DEVICE pointers
 * (stride in elements: 1 for a terminal-spot vector, ld for the last column of a path matrix -- pass S + n_steps).
 * Asynchronous on the handle's stream.  With b200mc_generate_paths / b200mc_simulate_terminal before and
 * b200mc_risk_metrics after, BASELINE config 4 (paths -> terminal P&L -> VaR/CVaR) never leaves the GPU. */
int b200mc_option_pnl(b200mc_handle *h, const void *S_dev, int64_t n, int64_t stride, int dtype_in, double strike,
                      int is_call, double discount, double premium, int dtype_out, void *pnl_dev);

/* Multi-rank form of the same metrics: every rank holds a shard of the P&L vector in its own HBM and
 * calls these three primitives; the host all-reduces their small outputs (2 doubles, 512 counters per radix pass,
 * 6 doubles) and takes the decisions, so every rank walks the same radix tree and ends with the GLOBAL order
 * statistics -- no P&L element ever leaves its GPU.  monte_carlo_option_simulator_b200/risk.py
 * (compute_risk_metrics_sharded) is the host side.  The state lives in the handle's scratch: do not interleave other
 * calls on the same handle between begin and finish.
 *   begin : out = { sum x, count of x < 0 } of the local shard (n may be 0)
 *   hist  : pass 7..0 (most significant byte first); hist[s][b] = number of local keys whose bytes above `pass` equal
 *           those of prefix[s] and whose byte `pass` is b; keys are the order-preserving 64-bit images of the doubles
 *           (negative: ~bits, else bits | 2^63)
 *   finish: out = { sum d^2, sum d^3, sum d^4, count(x < thr[0]), sum(x < thr[0]), sum log(x / thr[1]) over x < thr[1] },
 *           d = x - mean */
int b200mc_risk_begin(b200mc_handle *h, const void *pnl, int64_t n, int dtype, int on_device, double out[2]);
int b200mc_risk_hist(b200mc_handle *h, int pass, int nsel, const uint64_t prefix[2], uint64_t hist[512]);
int b200mc_risk_finish(b200mc_handle *h, double mean, int nsel, const double thr[2], double out[6]);

/* ---- draws, for feeding the reference the identical numbers ----------------------------------------------
 * out is float64 [n_paths, n_steps] on the host; `stream` one of B200MC_STREAM_*; `which` one of B200MC_Z*;
 * jump_prob = lambda_j * T / n_steps of the run to be reproduced (used only for B200MC_ZJUMP_SIZE of the SVJ stream).
 * Arrays a stream does not carry come back as neutral values (Z2 = 0, Z_jump = 1, Z_jump_size = 0). */
int b200mc_dump_normals(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths,
                        int32_t n_steps, uint32_t stream, int which, double jump_prob, double *out);
/* The stream a fused call with these parameters draws from: B200MC_STREAM_GBM when xi == 0 and lambda_j dt <= 0 (constant
 * or deterministic variance; deterministic-variance runs of more than 4096 steps fall to the Heston stream),
 * B200MC_STREAM_HESTON when lambda_j dt <= 0, else B200MC_STREAM_SVJ (always, with B200MC_FORCE_SVJ).  h may be NULL. */
int b200mc_select_stream(b200mc_handle *h, const b200mc_svj_params *p, double T, int32_t n_steps, uint32_t flags,
                         const b200mc_bumps *bumps, uint32_t *stream);
/* Raw Philox words uint32 [n_paths, n_blocks, 4] (host), for the bit-exact check against the oracle. */
int b200mc_dump_philox(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths,
                       int32_t n_blocks, uint32_t stream, uint32_t *out);

/* Raw moments of the normals the GBM stream produces for paths [path_offset, path_offset + n_paths) x n_blocks
 * Philox blocks (8 normals each), accumulated in fp64: out = { count, sum z, sum z^2, sum z^3, sum z^4, sum z_a z_b over
 * the two members of each Box-Muller pair }.  The statistical certificate of the generator (tools/normal_moments.py). */
int b200mc_normal_moments(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths, int32_t n_blocks,
                          double out[6]);
/* Counts of consecutive normals of the GBM stream on a 64 x 64 grid of equiprobable cells, out[64 a + b] with a, b =
 * floor(64 Phi(z)) of the first / second member: lag 0 = the two members of every Box-Muller word, lag 1 = the second
 * member of a word with the first member of the next word.  4 (lag 0) or 3 (lag 1) pairs per Philox block. */
int b200mc_normal_hist2d(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths, int32_t n_blocks, int lag,
                         uint64_t out[4096]);
/* lag | B200MC_HIST_WIDE: the same counts for the validation twin (B200MC_WIDE_RNG: 2 pairs per block; lag 1 = the second
 * member of the first pair with the first member of the second). */
#define B200MC_HIST_WIDE 0x100
/* lag | B200MC_HIST_R01: the same counts for round 1's field layout (radius = low 23 bits, angle = top 23 bits of the word,
 * 14 bits shared), which no pricing kernel uses any more: keeps the measurement that retired it reproducible. */
#define B200MC_HIST_R01 0x200

/* ---- device memory helpers for callers without a CUDA runtime of their own (ctypes) ---------------------- */
int b200mc_malloc(b200mc_handle *h, size_t bytes, void **dev_ptr);
int b200mc_free(b200mc_handle *h, void *dev_ptr);
int b200mc_memcpy_h2d(b200mc_handle *h, void *dst_dev, const void *src_host, size_t bytes);
int b200mc_memcpy_d2h(b200mc_handle *h, void *dst_host, const void *src_dev, size_t bytes);
int b200mc_malloc_host(b200mc_handle *h, size_t bytes, void **host_ptr); /* pinned */
int b200mc_free_host(b200mc_handle *h, void *host_ptr);
/* CUDA-event timing on the handle's stream: call begin, enqueue work, call end -> elapsed ms */
int b200mc_timer_begin(b200mc_handle *h);
int b200mc_timer_end(b200mc_handle *h, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200MC_H */
