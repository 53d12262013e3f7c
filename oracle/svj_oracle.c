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
