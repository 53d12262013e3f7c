// hedge.cu -- the daily delta-hedging walk of HedgingBacktest.run_backtest (engine/risk.py:278-316) for all scenarios
// at once: one thread per scenario, fp64, the reference's operation order.
//
//   per day:  delta = bs_delta(S, K, t_rem, r, q, sqrt(v0))        engine/monte_carlo.py:45-55   risk.py:283
//             trade = delta - hedge; cost = |trade| S (txn + slip) / 10000; cash -= trade S + cost   :286-290
//             S *= exp((r - q - v0/2) dt + sqrt(v0 dt) z);  t_rem -= dt                              :292-294,307-308
//   end:      pnl = cash + hedge S - payoff(S)                                                       :311-316
//
// z: either the caller's normals (Z[scenario][day], host array: the reference draws default_rng(seed).standard_normal()
// scenario by scenario, day by day -- the same stream as one (n_scenarios, n_days) draw), or Philox draws in registers:
// counter = (scenario_lo, scenario_hi, day / 8, B200MC_STREAM_HEDGE), the GBM layout of philox.cuh (word i of a block
// gives the normals of days 8j+2i and 8j+2i+1), z = BM_SCALE * (double)raw as b200mc_dump_normals returns them.
#include "prep.cuh"

namespace b200mc {

struct HedgeArgs {
    double S0, K, T, dt;
    double c1;          // r - q + sigma^2 / 2         (bs_delta's drift term)
    double sigma;       // sqrt(v0)
    double q;
    double gbm_drift;   // (r - q - v0 / 2) dt         :293
    double gbm_vol;     // sqrt(v0 dt)
    double cost_bps;    // txn_cost_bps + slippage_bps :287
    int64_t n;
    uint64_t scen0;
    int32_t n_days, is_call;
    PhiloxKey key;
};

template <bool GIVEN>
__global__ void __launch_bounds__(128)
k_hedge_walk(const __grid_constant__ HedgeArgs a, const double *__restrict__ Z, const double *__restrict__ premiums,
             double *__restrict__ pnl, double *__restrict__ cost_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const bool call = a.is_call != 0;
    double S = a.S0, cash = premiums ? premiums[i] : 0.0, hedge = 0.0, total = 0.0, t_rem = a.T;
    const uint64_t scen = a.scen0 + (uint64_t)i;
    double zb[8];
    for (int day = 0; day < a.n_days; ++day) {
        if (t_rem <= 0.0) break;                                                        // :279-280
        double z;
        if constexpr (GIVEN) {
            z = Z[(size_t)i * a.n_days + day];
        } else {
            if ((day & 7) == 0) {
                const U4 u = philox4x32_10((uint32_t)scen, (uint32_t)(scen >> 32), (uint32_t)(day >> 3),
                                           B200MC_STREAM_HEDGE, a.key);
                const BM2 b0 = box_muller_word(u.x), b1 = box_muller_word(u.y), b2 = box_muller_word(u.z),
                          b3 = box_muller_word(u.w);
                zb[0] = B200MC_BM_SCALE * (double)b0.rc; zb[1] = B200MC_BM_SCALE * (double)b0.rs;
                zb[2] = B200MC_BM_SCALE * (double)b1.rc; zb[3] = B200MC_BM_SCALE * (double)b1.rs;
                zb[4] = B200MC_BM_SCALE * (double)b2.rc; zb[5] = B200MC_BM_SCALE * (double)b2.rs;
                zb[6] = B200MC_BM_SCALE * (double)b3.rc; zb[7] = B200MC_BM_SCALE * (double)b3.rs;
            }
            z = zb[0];
#pragma unroll
            for (int t = 1; t < 8; ++t) z = (day & 7) == t ? zb[t] : z;
        }
        // bs_delta, monte_carlo.py:52-55 (t_rem > 0 here)
        const double d1 = __ddiv_rn(__dadd_rn(log(__ddiv_rn(S, a.K)), __dmul_rn(a.c1, t_rem)),
                                    __dmul_rn(a.sigma, sqrt(t_rem)));
        const double dq = exp(__dmul_rn(-a.q, t_rem));
        const double delta = call ? __dmul_rn(dq, normcdf(d1)) : __dmul_rn(dq, __dadd_rn(normcdf(d1), -1.0));
        const double trade = __dadd_rn(delta, -hedge);                                   // :286
        const double cost = __ddiv_rn(__dmul_rn(__dmul_rn(fabs(trade), S), a.cost_bps), 10000.0);   // :287
        total = __dadd_rn(total, cost);
        cash = __dadd_rn(cash, -__dadd_rn(__dmul_rn(trade, S), cost));                   // :289
        hedge = delta;
        S = __dmul_rn(S, exp(__dadd_rn(a.gbm_drift, __dmul_rn(a.gbm_vol, z))));          // :293
        t_rem = __dadd_rn(t_rem, -a.dt);                                                 // :308
    }
    const double payoff = call ? fmax(S - a.K, 0.0) : fmax(a.K - S, 0.0);                // :311-314
    pnl[i] = __dadd_rn(__dadd_rn(cash, __dmul_rn(hedge, S)), -payoff);                   // :316
    if (cost_out) cost_out[i] = total;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_hedge_walk(b200mc_handle *h, const b200mc_svj_params *p, double S0, double strike, double T,
                                 int is_call, int32_t n_days, int64_t n_scenarios, double cost_bps,
                                 const double *premiums, const double *Z, uint64_t seed, uint64_t scenario_offset,
                                 double *final_pnl, double *txn_cost)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!p) return fail(h, B200MC_EINVAL, "params is NULL");
    if (!final_pnl) return fail(h, B200MC_EINVAL, "final_pnl is NULL");
    if (n_days <= 0 || n_scenarios <= 0) return fail(h, B200MC_EINVAL, "n_days and n_scenarios must be positive");
    if (!(T > 0.0) || !isfinite(T) || !isfinite(S0) || !isfinite(strike))
        return fail(h, B200MC_EINVAL, "T must be positive and finite; S0 and strike finite");
    if (!(p->v0 >= 0.0)) return fail(h, B200MC_EINVAL, "v0 must be non-negative");
    B200MC_CUDA(h, cudaSetDevice(h->device));

    HedgeArgs a;
    memset(&a, 0, sizeof(a));
    a.S0 = S0; a.K = strike; a.T = T;
    a.dt = T / (double)n_days;                                         // risk.py:258
    a.sigma = sqrt(p->v0);                                             // :260
    a.c1 = p->r - p->q + 0.5 * (a.sigma * a.sigma);                    // monte_carlo.py:52
    a.q = p->q;
    a.gbm_drift = (p->r - p->q - 0.5 * p->v0) * a.dt;
    a.gbm_vol = sqrt(p->v0 * a.dt);
    a.cost_bps = cost_bps;
    a.n = n_scenarios;
    a.scen0 = scenario_offset;
    a.n_days = n_days;
    a.is_call = is_call ? 1 : 0;
    a.key = philox_make_key(seed);

    // device staging: [Z n*days][premiums n][pnl n][cost n]
    const size_t nz = Z ? (size_t)n_scenarios * n_days : 0, n = (size_t)n_scenarios;
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (nz + 3 * n) * 8));
    double *dZ = (double *)h->d_stage, *dP = dZ + nz, *dO = dP + n, *dC = dO + n;
    if (Z) B200MC_CUDA(h, cudaMemcpyAsync(dZ, Z, nz * 8, cudaMemcpyHostToDevice, h->stream));
    if (premiums) B200MC_CUDA(h, cudaMemcpyAsync(dP, premiums, n * 8, cudaMemcpyHostToDevice, h->stream));
    const unsigned grid = (unsigned)((n_scenarios + 127) / 128);
    if (Z) k_hedge_walk<true><<<grid, 128, 0, h->stream>>>(a, dZ, premiums ? dP : nullptr, dO, dC);
    else k_hedge_walk<false><<<grid, 128, 0, h->stream>>>(a, nullptr, premiums ? dP : nullptr, dO, dC);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(final_pnl, dO, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (txn_cost) B200MC_CUDA(h, cudaMemcpyAsync(txn_cost, dC, n * 8, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
