// prep.cuh -- host-side preparation shared by the fused entry points: mode selection, the constants of the
// recurrence (engine/monte_carlo.py:205-210) and, for deterministic-variance runs, the per-step tables.
#pragma once
#include <math.h>

#include <vector>

#include "sim.cuh"

namespace b200mc {

struct Prep {
    ModelArgs m;
    PhiloxKey key;
    int mode;
    int wld;                       // row length of the DETVAR tables (n_steps rounded up to 8)
    std::vector<double> wtab;      // [3][wld]  sqrt(v_s dt) BM_SCALE, zero padded
    std::vector<double> dtab;      // [wld]     cumulative drift of the primary state after step s
};

#define B200MC_DETVAR_MAX_STEPS 4096

// Jump process of a run with per-step jump probability p = lambda_j dt (engine/monte_carlo.py:233, "U < lambda_j dt"):
// returns whether jumps can fire at all and the constant of the geometric gap draw, 1 / lg2(1 - p) (philox.cuh).
inline int jump_setup(double p, double *inv_lg2_q)
{
    *inv_lg2_q = 0.0;                                   // p >= 1: every step jumps, gap = floor(lg2 U * 0) = 0
    if (!(p > 0.0)) return 0;
    if (p < 1.0) *inv_lg2_q = M_LN2 / log1p(-p);
    return 1;
}

inline int prepare(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                   int64_t n_paths, uint64_t seed, uint32_t flags, const b200mc_bumps *bumps, Prep &o)
{
    if (!p) return fail(h, B200MC_EINVAL, "params is NULL");
    if (n_steps <= 0) return fail(h, B200MC_EINVAL, "n_steps must be positive");
    if (n_paths <= 0) return fail(h, B200MC_EINVAL, "n_paths must be positive");
    if (!(T > 0.0) || !isfinite(T)) return fail(h, B200MC_EINVAL, "T must be positive and finite");
    if (!isfinite(S0)) return fail(h, B200MC_EINVAL, "S0 must be finite");
    if (!(fabs(p->rho) <= 1.0)) return fail(h, B200MC_EINVAL, "rho must lie in [-1, 1]");
    if ((flags & B200MC_GREEKS) && !bumps) return fail(h, B200MC_EINVAL, "B200MC_GREEKS needs bumps");

    ModelArgs &m = o.m;
    memset(&m, 0, sizeof(m));
    const double dt = T / (double)n_steps;                                  // :206
    const double sqrt_dt = sqrt(dt);                                        // :207
    const double k = exp(p->mu_j + 0.5 * p->sigma_j * p->sigma_j) - 1.0;    // :209
    const double drift_comp = p->r - p->q - p->lambda_j * k;                // :210
    m.S0 = S0; m.T = T; m.dt = dt;
    m.drift_dt = drift_comp * dt;
    m.half_dt = 0.5 * dt;
    m.sqrt_dt_s = sqrt_dt * B200MC_BM_SCALE;
    m.one_m_kdt = 1.0 - p->kappa * dt;
    m.kdt_theta = p->kappa * dt * p->theta;
    m.theta = p->theta;
    m.xi_sqrt_dt_s = p->xi * sqrt_dt * B200MC_BM_SCALE;
    m.rho = p->rho;
    m.crho = sqrt(1.0 - p->rho * p->rho);
    m.mu_j = p->mu_j;
    m.sigma_j = p->sigma_j;
    m.sigma_j_s = p->sigma_j * B200MC_BM_SCALE;
    m.jump_on = jump_setup(p->lambda_j * dt, &m.jump_inv_lg2q);             // :233
    m.v0[0] = p->v0;
    m.v0[1] = (flags & B200MC_GREEKS) ? bumps->v0_up : p->v0;
    m.v0[2] = (flags & B200MC_GREEKS) ? bumps->v0_dn : p->v0;
    o.key = philox_make_key(seed);

    const int nvar = (flags & B200MC_GREEKS) ? 3 : 1;
    bool constant_var = true;
    for (int r = 0; r < nvar; ++r)
        if (!(m.v0[r] >= 0.0) || (p->kappa != 0.0 && p->theta != m.v0[r])) constant_var = false;

    if ((flags & B200MC_FORCE_SVJ) || m.jump_on) o.mode = MODE_SVJ;
    else if (p->xi != 0.0) o.mode = MODE_HESTON;
    else if (constant_var) o.mode = MODE_GBM;
    else if (n_steps <= B200MC_DETVAR_MAX_STEPS) o.mode = MODE_DETVAR;
    else o.mode = MODE_HESTON;

    o.wld = (n_steps + 7) & ~7;
    if (o.mode == MODE_GBM) {
        for (int r = 0; r < 3; ++r) {
            m.step_drift[r] = (drift_comp - 0.5 * m.v0[r]) * dt;            // :229
            m.x_drift[r] = m.step_drift[r] * (double)n_steps;
            m.x_w[r] = sqrt(m.v0[r]) * sqrt_dt * B200MC_BM_SCALE;           // :224,226,230
        }
    } else if (o.mode == MODE_DETVAR) {
        o.wtab.assign((size_t)3 * o.wld, 0.0);
        o.dtab.assign((size_t)o.wld, 0.0);
        for (int r = 0; r < 3; ++r) {
            double v = m.v0[r], cum = 0.0;
            for (int s = 0; s < n_steps; ++s) {
                const double vp = v > 0.0 ? v : 0.0;                        // :223
                o.wtab[(size_t)r * o.wld + s] = sqrt(vp) * sqrt_dt * B200MC_BM_SCALE;
                cum += (drift_comp - 0.5 * vp) * dt;
                if (r == 0) o.dtab[s] = cum;
                v = vp + p->kappa * (p->theta - vp) * dt;                   // :237 with xi = 0
                v = v > 0.0 ? v : 0.0;                                      // :238
            }
            m.x_drift[r] = cum;
        }
    }
    return 0;
}

} // namespace b200mc
