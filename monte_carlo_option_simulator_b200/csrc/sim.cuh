// sim.cuh -- the per-path simulator shared by the fused kernels (sums, terminal values, path store).
//
// Restates the recurrence of _simulate_svj_paths_numba (engine/monte_carlo.py:205-241) in log space with
// the draws produced in registers by Philox4x32-10 (philox.cuh) instead of read from four host arrays:
//     v+ = max(v, 0)                                   :223
//     x += (r - q - lambda k - v+/2) dt + sqrt(v+) sqrt(dt) Z1 + jump      :226-236   (S = S0 exp(x))
//     v  = max(v+ + kappa (theta - v+) dt + xi sqrt(v+) sqrt(dt) (rho Z1 + sqrt(1-rho^2) Z2), 0)   :227,237-238
// One thread owns one path and carries NS "states" over the SAME draws (common random numbers):
//     state 0            the primary path
//     state 1 (ANTI)     its antithetic twin (-Z1, -Z2, U, -Zjs), monte_carlo.py:318-324
//     next two (GREEKS)  the path started from v0_up / v0_dn, greeks.py:124-147
// Four modes, picked by the host from the parameters:
//     GBM     xi = 0, lambda = 0, variance constant: x_T = drift + w * sum(z); only sum(z) is carried
//     DETVAR  xi = 0, lambda = 0, variance deterministic but moving (kappa != 0, theta != v0): per-step
//             weights w_s = sqrt(v_s dt) come from a table in shared memory (built on the host in fp64)
//     HESTON  lambda = 0: one Box-Muller pair (Z1, Z2) per step, four steps per Philox call
//     SVJ     everything: the same diffusion layout on its own stream; the jump times come from geometric gaps drawn
//             once per JUMP on a second stream (philox.cuh, JumpStream), not from a uniform per step
// R is the type of the path state (float or double); the draws are float in both cases.
#pragma once
#include "common.cuh"

// fp64 state of the constant-variance terminal kernels: 1 = a word's two normals enter as their exact fp32 sum (TwoSum,
// philox.cuh box_muller_word_twosum), 0 = each normal widened and added on its own
#ifndef B200MC_FP64_TWOSUM
#define B200MC_FP64_TWOSUM 0
#endif

namespace b200mc {

enum { MODE_GBM = 0, MODE_DETVAR = 1, MODE_HESTON = 2, MODE_SVJ = 3 };

struct ModelArgs {
    double S0, T, dt;
    double drift_dt;        // (r - q - lambda_j k) dt                       :210,229
    double half_dt;         // dt / 2
    double sqrt_dt_s;       // sqrt(dt) * BM_SCALE  (the draws are unscaled, see philox.cuh)
    double theta;
    double one_m_kdt;       // 1 - kappa dt
    double kdt_theta;       // kappa dt theta
    double jump_inv_lg2q;   // 1 / lg2(1 - lambda_j dt): gap to the next jump = floor(lg2 U * this)   (philox.cuh)
    double xi_sqrt_dt_s;    // xi sqrt(dt) BM_SCALE
    double rho, crho;       // crho = sqrt(1 - rho^2)                        :227
    double mu_j, sigma_j;
    double sigma_j_s;       // sigma_j BM_SCALE
    int32_t jump_on;        // lambda_j dt > 0                                :233
    int32_t pad_;
    double v0[3];           // base, up, down
    double x_drift[3];      // GBM / DETVAR: total drift of x over [0, T] for the three variance starts
    double x_w[3];          // GBM: x_T = x_drift + x_w * sum(raw z);  x_w = sqrt(v0 dt) BM_SCALE
    double step_drift[3];   // GBM: per-step drift (path store)
};

template <typename R> struct Consts {
    R drift_dt, half_dt, sqrt_dt_s, one_m_kdt, kdt_theta, xi_sqrt_dt_s, rho, crho, mu_j, sigma_j_s;
    R xs_rho, xs_crho;      // xi sqrt(dt) BM_SCALE * (rho, sqrt(1 - rho^2))
    __device__ __forceinline__ explicit Consts(const ModelArgs &m)
        : drift_dt((R)m.drift_dt), half_dt((R)m.half_dt), sqrt_dt_s((R)m.sqrt_dt_s), one_m_kdt((R)m.one_m_kdt),
          kdt_theta((R)m.kdt_theta), xi_sqrt_dt_s((R)m.xi_sqrt_dt_s), rho((R)m.rho), crho((R)m.crho),
          mu_j((R)m.mu_j), sigma_j_s((R)m.sigma_j_s), xs_rho((R)(m.xi_sqrt_dt_s * m.rho)),
          xs_crho((R)(m.xi_sqrt_dt_s * m.crho)) {}
};

__device__ __forceinline__ float  rsqrt_of(float x)  { return sqrt_approx(x); }
__device__ __forceinline__ double rsqrt_of(double x) { return sqrt(x); }
__device__ __forceinline__ float  rmax(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  rexp(float x)  { return expf(x); }
__device__ __forceinline__ double rexp(double x) { return exp(x); }

template <bool ANTI, bool GREEKS> struct StateLayout {
    static constexpr int NS = 1 + (ANTI ? 1 : 0) + (GREEKS ? 2 : 0);
    static constexpr int ANTI_IDX = 1;
    static constexpr int UP_IDX = 1 + (ANTI ? 1 : 0);
    static constexpr int DN_IDX = UP_IDX + 1;
};

// One stochastic-variance step for all states.  Lean form of monte_carlo.py:223-238:
//   * the constant drift (r - q - lambda k) dt is NOT added here: n_steps * drift_dt is added once at the end;
//   * v is carried unclamped -- the clamp of :238 is the max(v, 0) of :223 at the next step (and of the final v_T);
//   * jumps are added by the caller inside the (rare) branch that detects them.
// The draws of the step arrive as StepDraws (scaled_draws below), prepared ONCE per step, not once per state:
//   fp64 state   a = sqrt(dt) Z1, b = xi sqrt(dt) (rho Z1 + sqrt(1-rho^2) Z2), and every state takes q = sqrt(v+);
//   fp32 state   both noise terms of :230,237 carry sqrt(v+) times the Box-Muller radius sqrt(-2 ln u), and
//                sqrt(v+) sqrt(L) = sqrt(v+ L): the radius is never formed -- a = sqrt(dt) cos, b = xi sqrt(dt)
//                (rho cos + sqrt(1-rho^2) sin), L = -lg2 u (constants folded), q = sqrt(v+ L).  One MUFU less per step
//                than radius + sqrt(v+) (5 instead of 6 for an antithetic pair), the same FP32 count.
// 5 FP32 + 1 MUFU per state and step.
template <typename R> struct StepDraws { R a, b, L; };

template <typename R, bool ANTI, bool GREEKS>
__device__ __forceinline__ void sv_step(R (&x)[StateLayout<ANTI, GREEKS>::NS], R (&v)[StateLayout<ANTI, GREEKS>::NS],
                                        const Consts<R> &c, const StepDraws<R> &d)
{
    constexpr int NS = StateLayout<ANTI, GREEKS>::NS;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const bool neg = ANTI && k == 1;                 // twin: -Z1, -Z2  (:323)
        const R vp = rmax(v[k], (R)0);                   // :223
        R sv;
        if constexpr (sizeof(R) == 4) sv = rsqrt_of(vp * d.L);
        else sv = rsqrt_of(vp);                          // :224
        const R t = x[k] - c.half_dt * vp;               // :229 without the constant part
        const R mr = vp * c.one_m_kdt + c.kdt_theta;     // v+ + kappa (theta - v+) dt      :237
        x[k] = neg ? t - sv * d.a : t + sv * d.a;        // :230,236
        v[k] = neg ? mr - sv * d.b : mr + sv * d.b;      // :237
    }
}

// The draws of a step from its Box-Muller word.  fp64 state: from the fp32 pair exactly as b200mc_dump_normals exports
// it (Z = BM_SCALE * (double)raw), so the oracle fed the dumped draws agrees to rounding.  fp32 state: see sv_step (the
// results differ from the fp64 route by fp32 rounding only, far inside the 1e-4 band).
template <typename R>
__device__ __forceinline__ StepDraws<R> scaled_draws(uint32_t w, const Consts<R> &c)
{
    StepDraws<R> d;
    if constexpr (sizeof(R) == 4) {
        const BM3L b = box_muller_parts_l(w);
        d.L = b.L;
        d.a = c.sqrt_dt_s * b.cs;
        d.b = fmaf(c.xs_crho, b.sn, c.xs_rho * b.cs);
    } else {
        const BM2 b = box_muller_word(w);
        d.L = (R)1;
        d.a = c.sqrt_dt_s * (R)b.rc;
        d.b = c.xi_sqrt_dt_s * (c.rho * (R)b.rc + c.crho * (R)b.rs);
    }
    return d;
}

// Draw layout (part of the ABI, include/b200mc.h "Random numbers"): Philox counter = (path_lo, path_hi, block, stream),
// every 32-bit output word w yields one Box-Muller pair (rc, rs) = box_muller_word(w):
//   GBM / DETVAR  stream 0, block j -> steps 8j..8j+7: word i gives the normals of steps 8j+2i (rc) and 8j+2i+1 (rs)
//   HESTON        stream 1, block j -> steps 4j..4j+3: word i gives (Z1, Z2) = (rc, rs) of step 4j+i
//   SVJ           stream 2, block j -> steps 4j..4j+3 like HESTON; jumps: stream 4, block k -> jumps 2k, 2k+1 of the path,
//                 (w0, w2) -> the geometric gaps before them, (w1, w3) -> their sizes (philox.cuh, JumpStream)
//
// Simulates global path `path` to T.  On return xT[k] = log(S_T / S0) of state k and vT[k] its variance
// (GBM / DETVAR: vT is left untouched); sumz_out = sum of the raw draws (GBM only; feeds the pathwise vega).
// wtab: DETVAR weights, wtab[k * wld + s] = sqrt(v_s^{(k)} dt) * BM_SCALE in shared memory (wld = steps rounded up to 8,
// zero padded).  rec: after every step s, rec(s, x0) gets the primary state's log return (path store).
// PAIRED (GBM): two Philox blocks per loop iteration (see the GBM branch); chosen by the caller -- it pays in the
// single-strike fp32 kernel (+3.6 %) and costs 1.5-3 % where the loop is XU bound (fp64 state) or short of registers (multi-strike).
template <int MODE, bool ANTI, bool GREEKS, typename R, typename Rec, bool WIDE = false, bool PAIRED = false>
__device__ __forceinline__ void simulate_path(const ModelArgs &m, const PhiloxKey &key, uint64_t path, int n_steps,
                                              const R *wtab, int wld,
                                              R (&xT)[StateLayout<ANTI, GREEKS>::NS],
                                              R (&vT)[StateLayout<ANTI, GREEKS>::NS], R &sumz_out, Rec rec)
{
    using L = StateLayout<ANTI, GREEKS>;
    constexpr int NS = L::NS;
    const uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32);

    if constexpr (MODE == MODE_GBM && Rec::enabled) {
        // path store: the log return itself is carried so that every step can be recorded
        const R w = (R)m.x_w[0], d = (R)m.step_drift[0];
        R x = (R)0;
        const int nblk = (n_steps + 7) >> 3;
        for (int j = 0; j < nblk; ++j) {
            const U4 u = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
            const BM2 b0 = box_muller_word(u.x), b1 = box_muller_word(u.y), b2 = box_muller_word(u.z),
                      b3 = box_muller_word(u.w);
            const R z[8] = {(R)b0.rc, (R)b0.rs, (R)b1.rc, (R)b1.rs, (R)b2.rc, (R)b2.rs, (R)b3.rc, (R)b3.rs};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (8 * j + t < n_steps) {
                    x += w * z[t] + d;
                    rec(8 * j + t, x);
                }
            }
        }
        sumz_out = (R)0;
        xT[0] = x;
    } else if constexpr (MODE == MODE_GBM && WIDE) {
        // validation twin: one Box-Muller pair per TWO words (philox.cuh, box_muller_wide), block j -> steps 4j..4j+3
        R sumz = (R)0;
        const int nb = (n_steps + 3) >> 2;
        for (int j = 0; j < nb; ++j) {
            const U4 u = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
            const BM2 b0 = box_muller_wide(u.x, u.y), b1 = box_muller_wide(u.z, u.w);
            const R z[4] = {(R)b0.rc, (R)b0.rs, (R)b1.rc, (R)b1.rs};
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (4 * j + t < n_steps) sumz += z[t];
        }
        sumz_out = sumz;
        xT[0] = (R)m.x_drift[0] + (R)m.x_w[0] * sumz;
        if constexpr (ANTI) xT[1] = (R)m.x_drift[0] - (R)m.x_w[0] * sumz;
        if constexpr (GREEKS) {
            xT[L::UP_IDX] = (R)m.x_drift[1] + (R)m.x_w[1] * sumz;
            xT[L::DN_IDX] = (R)m.x_drift[2] + (R)m.x_w[2] * sumz;
        }
    } else if constexpr (MODE == MODE_GBM) {
        // Software-pipelined: the Philox rounds of block j+1 (IMAD.WIDE / LOP3) are issued in the same loop body as
        // the Box-Muller transforms of block j (MUFU), so every warp feeds the XU pipe at an even rate instead of
        // in bursts (ncu: mio_throttle was the top stall with the two phases back to back).
        // fp32 state: only the SUM of the draws enters x_T, and the two normals of a word add up to sqrt(2) rad sin(a + pi/4)
        // (philox.cuh, box_muller_word_sum): full blocks take 3 MUFU and one accumulation per word instead of 4 and two.
        // The fp64 state adds every widened normal on its own, as b200mc_dump_normals exports them (parity to rounding).
        R sumz = (R)0;
        float sumw = 0.f;                                  // fp32 state: sum over whole words of rad sin(a + pi/4);
                                                           // fp64 state (B200MC_FP64_TWOSUM): the words' TwoSum remainders
        const int nb = (n_steps + 7) >> 3;                 // blocks, the last one possibly partial
        auto whole = [&](const U4 &q) {                    // all eight draws of a block
            if constexpr (sizeof(R) == 4) {
                sumw += box_muller_word_sum(q.x);
                sumw += box_muller_word_sum(q.y);
                sumw += box_muller_word_sum(q.z);
                sumw += box_muller_word_sum(q.w);
            } else {
#if B200MC_FP64_TWOSUM
                // the exact fp32 sum of a word's two normals, widened once; the exact remainders go to an fp32 side sum
                const BMSum b0 = box_muller_word_twosum(q.x), b1 = box_muller_word_twosum(q.y),
                            b2 = box_muller_word_twosum(q.z), b3 = box_muller_word_twosum(q.w);
                sumz += (R)b0.s; sumz += (R)b1.s; sumz += (R)b2.s; sumz += (R)b3.s;
                sumw += (b0.e + b1.e) + (b2.e + b3.e);
#else
                const BM2 b0 = box_muller_word(q.x), b1 = box_muller_word(q.y), b2 = box_muller_word(q.z),
                          b3 = box_muller_word(q.w);
                sumz += (R)b0.rc; sumz += (R)b0.rs; sumz += (R)b1.rc; sumz += (R)b1.rs;
                sumz += (R)b2.rc; sumz += (R)b2.rs; sumz += (R)b3.rc; sumz += (R)b3.rs;
#endif
            }
        };
        // two blocks per iteration, the two word sets alternating between u and un: no register copies at the loop end
        // (they were 5 of the 85 instructions of an iteration, and the issue port is what this loop is bound by)
        U4 u = philox4x32_10(c0, c1, 0u, B200MC_STREAM_GBM, key);
        if constexpr (PAIRED) {
            int j = 1;
            for (; j + 1 < nb; j += 2) {
                const U4 un = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
                whole(u);
                u = philox4x32_10(c0, c1, (uint32_t)(j + 1), B200MC_STREAM_GBM, key);
                whole(un);
            }
            if (j < nb) {
                const U4 un = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
                whole(u);
                u = un;
            }
        } else {
            for (int j = 1; j < nb; ++j) {
                const U4 un = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
                whole(u);
                u = un;
            }
        }
        {
            const int rem = n_steps - 8 * (nb - 1);        // 1..8 draws of the last block are used
            const BM2 b0 = box_muller_word(u.x), b1 = box_muller_word(u.y), b2 = box_muller_word(u.z),
                      b3 = box_muller_word(u.w);
            const R z[8] = {(R)b0.rc, (R)b0.rs, (R)b1.rc, (R)b1.rs, (R)b2.rc, (R)b2.rs, (R)b3.rc, (R)b3.rs};
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < rem) sumz += z[t];
        }
        if constexpr (sizeof(R) == 4) sumz = fmaf(sumw, B200MC_SQRT2_F, sumz);
#if B200MC_FP64_TWOSUM
        else sumz += (R)sumw;                              // the words' exact remainders (fp64 state)
#endif
        sumz_out = sumz;
        xT[0] = (R)m.x_drift[0] + (R)m.x_w[0] * sumz;
        if constexpr (ANTI) xT[1] = (R)m.x_drift[0] - (R)m.x_w[0] * sumz;
        if constexpr (GREEKS) {
            xT[L::UP_IDX] = (R)m.x_drift[1] + (R)m.x_w[1] * sumz;
            xT[L::DN_IDX] = (R)m.x_drift[2] + (R)m.x_w[2] * sumz;
        }
    } else if constexpr (MODE == MODE_DETVAR) {
        R acc[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) acc[k] = (R)0;
        // weight row of state k (the ANTI twin shares row 0 with a minus sign and is not accumulated)
        auto row = [&](int k) -> int { return k == 0 ? 0 : k - (ANTI ? 1 : 0); };
        const int nblk = (n_steps + 7) >> 3;     // the table is zero-padded to a multiple of 8
        for (int j = 0; j < nblk; ++j) {
            const U4 u = philox4x32_10(c0, c1, (uint32_t)j, B200MC_STREAM_GBM, key);
            const BM2 b0 = box_muller_word(u.x), b1 = box_muller_word(u.y), b2 = box_muller_word(u.z),
                      b3 = box_muller_word(u.w);
            const R z[8] = {(R)b0.rc, (R)b0.rs, (R)b1.rc, (R)b1.rs, (R)b2.rc, (R)b2.rs, (R)b3.rc, (R)b3.rs};
            // the 8 weights of a row come in with 128-bit shared loads (rows are 16-byte aligned, wld % 8 == 0): every
            // scalar LDS would be one more entry in the MIO queue the MUFUs already saturate
            R w[NS][8];
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                if (ANTI && k == 1) continue;
                const int4 *src = reinterpret_cast<const int4 *>(wtab + row(k) * wld + 8 * j);
                int4 *dst = reinterpret_cast<int4 *>(w[k]);
#pragma unroll
                for (int q = 0; q < (int)(8 * sizeof(R) / 16); ++q) dst[q] = src[q];
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    if (ANTI && k == 1) continue;
                    acc[k] += w[k][t] * z[t];
                }
                if constexpr (Rec::enabled) { if (8 * j + t < n_steps) rec(8 * j + t, acc[0]); }
            }
        }
        sumz_out = (R)0;
        xT[0] = (R)m.x_drift[0] + acc[0];
        if constexpr (ANTI) xT[1] = (R)m.x_drift[0] - acc[0];
        if constexpr (GREEKS) {
            xT[L::UP_IDX] = (R)m.x_drift[1] + acc[L::UP_IDX];
            xT[L::DN_IDX] = (R)m.x_drift[2] + acc[L::DN_IDX];
        }
    } else {
        const Consts<R> c(m);
        R x[NS], v[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) { x[k] = (R)0; v[k] = (R)m.v0[0]; }
        if constexpr (GREEKS) { v[L::UP_IDX] = (R)m.v0[1]; v[L::DN_IDX] = (R)m.v0[2]; }
        R dacc = (R)0;                                   // running constant drift, only for the path store
        // Software-pipelined like the GBM loop: the Philox rounds of block j+1 are issued next to the transforms and
        // variance steps of block j.
        JumpStream jmp;
        const float inv_lg2_q = (float)m.jump_inv_lg2q;
        if constexpr (MODE == MODE_SVJ) jmp.init(c0, c1, key, inv_lg2_q, m.jump_on != 0);
        // Terminal values only (nothing records the path): the jumps add to the log spot and touch nothing else -- the
        // variance recurrence never sees them (monte_carlo.py:229-238) -- so their total, n_J mu_J + sigma_J sum Z_i, is
        // accumulated HERE, before the step loop, by walking the path's jump times.  Inside the step loop a jump is a
        // divergent branch that costs the whole warp ~100 cycles for ONE lane (12 % of the warp-steps at lambda dt = 0.004:
        // 12 of the 54 instructions per pair-step); here every lane walks its own jumps at the same time, with the refills
        // of all lanes in the same iteration: max-over-lanes(n_J) short iterations per path instead of one divergent
        // excursion per jump and lane.  The draws are unchanged; only the order of the additions into x differs.
        R jump_mu = (R)0, jump_z = (R)0;
        if constexpr (MODE == MODE_SVJ && !Rec::enabled) {
            while (jmp.next < n_steps) {                                                 // :233-234
                jump_z += (R)jmp.size_raw();
                jump_mu += c.mu_j;
                jmp.advance(c0, c1, key, inv_lg2_q, jmp.next);
            }
        }
        constexpr uint32_t STREAM = MODE == MODE_HESTON ? B200MC_STREAM_HESTON : B200MC_STREAM_SVJ;
        // one step from its word (the recorder and, for a recorded SVJ path, the in-place jump ride along)
        auto step = [&](uint32_t w, int s) {
            const StepDraws<R> dr = scaled_draws<R>(w, c);
            sv_step<R, ANTI, GREEKS>(x, v, c, dr);
            if constexpr (MODE == MODE_SVJ && Rec::enabled) {                     // a recorded path takes its jumps in place
                if (s == jmp.next) {                                                     // :233-234, rare
                    const R jsz = c.sigma_j_s * (R)jmp.size_raw();
#pragma unroll
                    for (int k = 0; k < NS; ++k) x[k] += (ANTI && k == 1) ? c.mu_j - jsz : c.mu_j + jsz;
                    jmp.advance(c0, c1, key, inv_lg2_q, s);
                }
            }
            if constexpr (Rec::enabled) { dacc += c.drift_dt; rec(s, x[0] + dacc); }
        };
        // Full blocks of four steps run without a per-step bound test (three compares and branches per block: 4 % of the
        // loop); the last 1-3 steps of a path are a separate tail.  The Philox rounds of block j + 1 are issued before the
        // steps of block j, as in the GBM loop.
        const int nfull = n_steps >> 2, rem = n_steps & 3, nblk = nfull + (rem ? 1 : 0);
        U4 u = philox4x32_10(c0, c1, 0u, STREAM, key);
        for (int j = 0; j < nfull; ++j) {
            U4 un = u;
            if (j + 1 < nblk) un = philox4x32_10(c0, c1, (uint32_t)(j + 1), STREAM, key);
            step(u.x, 4 * j);
            step(u.y, 4 * j + 1);
            step(u.z, 4 * j + 2);
            step(u.w, 4 * j + 3);
            u = un;
        }
        if (rem) {
            step(u.x, 4 * nfull);
            if (rem > 1) step(u.y, 4 * nfull + 1);
            if (rem > 2) step(u.z, 4 * nfull + 2);
        }
        {
            const R total_drift = (R)((double)n_steps * m.drift_dt);
            const R jsz = c.sigma_j_s * jump_z;                                          // 0 unless SVJ terminal mode
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                x[k] += total_drift + ((ANTI && k == 1) ? jump_mu - jsz : jump_mu + jsz);
                v[k] = rmax(v[k], (R)0);
            }
        }
        sumz_out = (R)0;
#pragma unroll
        for (int k = 0; k < NS; ++k) { xT[k] = x[k]; vT[k] = v[k]; }
    }
}

struct NoRec {
    static constexpr bool enabled = false;
    template <typename R> __device__ __forceinline__ void operator()(int, R) const {}
};

} // namespace b200mc
