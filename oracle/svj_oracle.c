/*
 * oracle/svj_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The shipped path is the CUDA
 * library (libb200mc.so) and it never links or calls anything in here.
 *
 * Parity status: PINNED.  The reference is Python and imports in the dev container, so every
 * function here is checked (tests/test_oracle.py) against fixtures produced by running the
 * reference itself (tests/golden/make_golden.py -> tests/golden/ npz + json fixtures).
 *
 * What is restated (citations are file:line in the reference tree):
 *   oracle_simulate_svj     engine/monte_carlo.py:189-243  (_simulate_svj_paths_numba)
 *   oracle_philox4x32_10    not in the reference: Random123 Philox4x32-10, the counter-based
 *                           generator the CUDA path uses (KATs in SURVEY.md section 8c).
 *
 * The reference iterates step-outer / path-inner (monte_carlo.py:221-222).  Paths never interact,
 * so the loop nest is swapped here (path-outer) -- per path the floating-point operation order is
 * the same as the reference's, hence results are identical up to libm exp/sqrt rounding.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC; no -ffast-math: the operation order
 * is part of the contract).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* engine/monte_carlo.py:189-243.  Arrays are C-contiguous [n_paths, n_steps] float64.
 * all_paths (may be NULL when record_paths == 0) is [n_paths, n_steps + 1], column 0 = S0 (:216-217). */
void oracle_simulate_svj(double S0, double v0, double r, double q, double T,
                         double kappa, double theta, double xi, double rho,
                         double lambda_j, double mu_j, double sigma_j,
                         const double *Z1, const double *Z2,
                         const double *Z_jump, const double *Z_jump_size,
                         int64_t n_paths, int32_t n_steps, int record_paths,
                         double *S_final, double *v_final, double *all_paths)
{
    const double dt = T / (double)n_steps;                         /* :206 */
    const double sqrt_dt = sqrt(dt);                               /* :207 */
    const double k = exp(mu_j + 0.5 * (sigma_j * sigma_j)) - 1.0;  /* :209 */
    const double drift_comp = r - q - lambda_j * k;                /* :210 */
    const double sq1mr2 = sqrt(1.0 - rho * rho);                   /* :227 */
    const double jump_thr = lambda_j * dt;                         /* :233 */

#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_paths; ++i) {
        double S = S0, v = v0;                                     /* :212-213 */
        const size_t row = (size_t)i * (size_t)n_steps;
        double *prow = NULL;
        if (record_paths) {
            prow = all_paths + (size_t)i * (size_t)(n_steps + 1);
            prow[0] = S0;                                          /* :217 */
        }
        for (int32_t s = 0; s < n_steps; ++s) {
            const double v_pos = v > 0.0 ? v : 0.0;                /* :223 */
            const double sqrt_v = sqrt(v_pos);                     /* :224 */
            const double z1 = Z1[row + s];
            const double dW1 = z1 * sqrt_dt;                       /* :226 */
            const double dW2 = rho * z1 * sqrt_dt + sq1mr2 * Z2[row + s] * sqrt_dt; /* :227 */
            const double log_drift = (drift_comp - 0.5 * v_pos) * dt;               /* :229 */
            const double log_diffusion = sqrt_v * dW1;             /* :230 */
            double jump = 0.0;                                     /* :232 */
            if (Z_jump[row + s] < jump_thr)                        /* :233 */
                jump = mu_j + sigma_j * Z_jump_size[row + s];      /* :234 */
            S = S * exp(log_drift + log_diffusion + jump);         /* :236 */
            v = v_pos + kappa * (theta - v_pos) * dt + xi * sqrt_v * dW2; /* :237 */
            v = v > 0.0 ? v : 0.0;                                 /* :238 */
            if (record_paths) prow[s + 1] = S;                     /* :241 */
        }
        S_final[i] = S;
        v_final[i] = v;
    }
}

/* Random123 Philox4x32-10 (Salmon et al., SC'11): 10 rounds of two 32x32->64 multiplies and a
 * Feistel-style swap, key bumped by the Weyl constants between rounds. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Raw words for a rectangle of counters, in the layout the CUDA path defines (include/b200mc.h):
 * ctr = (path_lo, path_hi, block, stream), key = (seed_lo, seed_hi).  out is [n_paths, n_blocks, 4]. */
void oracle_philox_block_words(uint64_t seed, uint64_t path_offset, int64_t n_paths,
                               int32_t n_blocks, uint32_t stream, uint32_t *out)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_paths; ++i) {
        const uint64_t p = path_offset + (uint64_t)i;
        for (int32_t b = 0; b < n_blocks; ++b) {
            const uint32_t ctr[4] = { (uint32_t)p, (uint32_t)(p >> 32), (uint32_t)b, stream };
            oracle_philox4x32_10(ctr, key, out + ((size_t)i * (size_t)n_blocks + (size_t)b) * 4);
        }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------------------
 * NumPy's pseudo-random front end (a third-party dependency of the reference: numpy 2.3.5 here; the reference calls
 * np.random.default_rng(seed).standard_normal((n, steps)) x 3 and .random((n, steps)), engine/monte_carlo.py:301-308,
 * engine/greeks.py:33-41).  Restated from the published algorithm -- PCG64 (pcg64 XSL-RR 128/64) and the 256-layer
 * Ziggurat of numpy/random/src/distributions/distributions.c (random_standard_normal) -- as the CPU checker of
 * csrc/np_normal.cu.  Pinned: tests/test_oracle.py compares it bit for bit with NumPy itself in the dev container.
 * The three tables (ki_double / wi_double / fi_double) are passed in by the caller. */
static inline uint64_t pcg64_next(unsigned __int128 *s, unsigned __int128 inc)
{
    const unsigned __int128 MULT = ((unsigned __int128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    *s = *s * MULT + inc;
    const uint64_t hi = (uint64_t)(*s >> 64), lo = (uint64_t)*s, x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
static inline double pcg64_double(uint64_t o) { return (double)(o >> 11) * (1.0 / 9007199254740992.0); }

/* state = {state_hi, state_lo, inc_hi, inc_lo} BEFORE the first output wanted.  Returns the generator outputs consumed. */
uint64_t oracle_np_standard_normal(const uint64_t state[4], const uint64_t *ki, const double *wi, const double *fi,
                                   int64_t n, double *out)
{
    const double R = 3.6541528853610087963519472518, INV_R = 0.27366123732975827203338247596;
    unsigned __int128 s = ((unsigned __int128)state[0] << 64) | state[1], inc = ((unsigned __int128)state[2] << 64) | state[3];
    uint64_t used = 0;
    for (int64_t i = 0; i < n;) {
        uint64_t r = pcg64_next(&s, inc);
        ++used;
        const int idx = (int)(r & 0xff);
        r >>= 8;
        const int sign = (int)(r & 1);
        const uint64_t rabs = (r >> 1) & 0x000fffffffffffffull;
        double x = (double)rabs * wi[idx];
        if (sign) x = -x;
        if (rabs < ki[idx]) { out[i++] = x; continue; }
        if (idx == 0) {
            for (;;) {
                const double xx = -INV_R * log1p(-pcg64_double(pcg64_next(&s, inc)));
                const double yy = -log1p(-pcg64_double(pcg64_next(&s, inc)));
                used += 2;
                if (yy + yy > xx * xx) { out[i++] = ((rabs >> 8) & 1) ? -(R + xx) : R + xx; break; }
            }
        } else {
            const double u = pcg64_double(pcg64_next(&s, inc));
            ++used;
            if ((fi[idx - 1] - fi[idx]) * u + fi[idx] < exp(-0.5 * x * x)) out[i++] = x;
        }
    }
    return used;
}

/* glibc 2.28+ sysdeps/ieee754/dbl-64/s_log1p.c in the operation order and with the fused multiply-adds of its x86-64 FMA
 * build (the variant the dynamic loader picks on hosts with FMA, hence what NumPy's log1p is there).  csrc/np_normal.cu
 * carries the same sequence for the tail draws of the Ziggurat; tests/test_oracle.py checks this function bit for bit
 * against the host's log1p.  (File built with -ffp-contract=off: only the explicit fma() calls fuse.) */
double oracle_glibc_log1p_fma(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    union { double d; uint64_t u; } b;
    double f = 0.0, c = 0.0, u;
    int32_t hu = 0, k = 1;
    b.d = x;
    const int32_t hx = (int32_t)(b.u >> 32), ax = hx & 0x7fffffff;
    if (hx < 0x3FDA827A) {
        if (ax >= 0x3ff00000) return x == -1.0 ? -INFINITY : NAN;
        if (ax < 0x3e200000) return ax < 0x3c900000 ? x : fma(-(x * x), 0.5, x);
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) { k = 0; f = x; hu = 1; }
    }
    if (hx >= 0x7ff00000) return x + x;
    if (k != 0) {
        if (hx < 0x43400000) {
            u = 1.0 + x;
            b.d = u; hu = (int32_t)(b.u >> 32);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0);
            c /= u;
        } else {
            u = x;
            b.d = u; hu = (int32_t)(b.u >> 32);
            k = (hu >> 20) - 1023;
            c = 0;
        }
        hu &= 0x000fffff;
        b.d = u;
        if (hu < 0x6a09e) {
            b.u = (b.u & 0xffffffffull) | ((uint64_t)(uint32_t)(hu | 0x3ff00000) << 32);
        } else {
            k += 1;
            b.u = (b.u & 0xffffffffull) | ((uint64_t)(uint32_t)(hu | 0x3fe00000) << 32);
            hu = (0x00100000 - hu) >> 2;
        }
        u = b.d;
        f = u - 1.0;
    }
    const double hfsq = 0.5 * f * f, dk = (double)k;
    if (hu == 0) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = fma(dk, ln2_lo, c);
            return fma(dk, ln2_hi, c);
        }
        const double R = hfsq * fma(-0.66666666666666666, f, 1.0);
        if (k == 0) return f - R;
        return fma(dk, ln2_hi, -((R - fma(dk, ln2_lo, c)) - f));
    }
    const double s = f / (2.0 + f), z = s * s;
    const double z2 = z * z, R2 = fma(z, Lp3, Lp2), z4 = z2 * z2, R3 = fma(z, Lp5, Lp4), z6 = z4 * z2, R4 = fma(z, Lp7, Lp6);
    const double R = fma(z6, R4, fma(z4, R3, fma(z, Lp1, z2 * R2)));
    const double sp = s * (hfsq + R);
    if (k == 0) return f - (hfsq - sp);
    return fma(dk, ln2_hi, -((hfsq - (sp + fma(dk, ln2_lo, c))) - f));
}

/* bitwise mismatches between oracle_glibc_log1p_fma and the host's log1p over n arguments -U, U uniform in [0, 1),
 * and over the same arguments scaled down by random powers of two (the small-|x| branches) */
int64_t oracle_log1p_mismatches(int64_t n, uint64_t seed)
{
    uint64_t st = seed ? seed : 88172645463325252ull;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        const double u = (double)(st >> 11) * (1.0 / 9007199254740992.0);
        const double a = -u, b = -u * ldexp(1.0, -(int)(st & 63));
        union { double d; uint64_t u; } p, q;
        p.d = log1p(a); q.d = oracle_glibc_log1p_fma(a); bad += p.u != q.u;
        p.d = log1p(b); q.d = oracle_glibc_log1p_fma(b); bad += p.u != q.u;
    }
    return bad;
}
