/* examples/price_c.c -- libb200mc from plain C: no Python, no torch, only include/b200mc.h.
 *
 *   gcc -O2 -Iinclude examples/price_c.c -o /tmp/price_c -Lmonte_carlo_option_simulator_b200 -lb200mc \
 *       -Wl,-rpath,$PWD/monte_carlo_option_simulator_b200 -lm
 *   /tmp/price_c [n_paths]
 *
 * Prices a European call at three strikes over shared paths (what MonteCarloEngine.price_batch does,
 * engine/monte_carlo.py:377-450), computes VaR / CVaR of the terminal P&L (engine/risk.py:117-155) and inverts the
 * prices to implied volatilities (engine/surface.py:48-66).  Prints one line per result; exits non-zero with the
 * library's message when there is no sm_100 device (there is no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200mc.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != 0) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, b200mc_last_error(h));   \
            if (h) b200mc_destroy(h);                                                    \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv)
{
    b200mc_handle *h = NULL;
    const int64_t n_paths = argc > 1 ? atoll(argv[1]) : 1000000;
    CHECK(b200mc_create(0, &h));

    /* the reference's default SVJ parameters (engine/models.py:31-44) */
    b200mc_svj_params p = {0.04, 0.065, 0.012, 3.0, 0.04, 0.5, -0.7, 1.0, -0.05, 0.10};
    const double S0 = 22500.0, T = 0.25, strikes[3] = {21000.0, 22500.0, 24000.0};
    const int32_t n_steps = 63;                     /* max(int(252 * T), 10), engine/monte_carlo.py:287 */
    b200mc_sums sums[3];
    CHECK(b200mc_price_european(h, &p, S0, T, n_steps, n_paths, 42, 0, strikes, 3, 1, B200MC_ANTITHETIC, NULL, sums));
    double prices[3], mats[3] = {T, T, T}, iv[3];
    int32_t is_call[3] = {1, 1, 1};
    for (int k = 0; k < 3; ++k) {
        const double mean = 0.5 * (sums[k].sum_a + sums[k].sum_b) / sums[k].n;      /* antithetic pair average */
        const double var = 0.25 * (sums[k].sum_aa + 2.0 * sums[k].sum_ab + sums[k].sum_bb) / sums[k].n - mean * mean;
        prices[k] = exp(-p.r * T) * mean;
        printf("K=%.0f price=%.6f std_error=%.6f\n", strikes[k], prices[k], exp(-p.r * T) * sqrt(var / sums[k].n));
    }
    CHECK(b200mc_implied_vol(h, 3, prices, strikes, mats, is_call, S0, p.r, p.q, 0.001, 5.0, iv));
    printf("iv=%.6f %.6f %.6f\n", iv[0], iv[1], iv[2]);

    /* terminal spots -> discounted P&L of the ATM call bought at its Monte Carlo price -> tail metrics */
    double *S = (double *)malloc((size_t)n_paths * sizeof(double));
    if (!S) return 3;
    CHECK(b200mc_simulate_terminal(h, &p, S0, T, n_steps, n_paths, 42, 0, B200MC_FP64, B200MC_F64, 0, S, NULL, NULL));
    for (int64_t i = 0; i < n_paths; ++i) S[i] = exp(-p.r * T) * fmax(S[i] - strikes[1], 0.0) - prices[1];
    double m[8];
    CHECK(b200mc_risk_metrics(h, S, n_paths, B200MC_F64, 0, 0.99, m));
    printf("var99=%.6f cvar99=%.6f skew=%.4f kurt=%.4f mean=%.6f\n", m[0], m[1], m[2], m[3], m[6]);
    free(S);
    printf("kernels launched: %lld\n", (long long)b200mc_launch_count(h));
    b200mc_destroy(h);
    return 0;
}
