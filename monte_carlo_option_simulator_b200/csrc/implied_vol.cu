// implied_vol.cu -- Black-Scholes implied volatilities for a whole option chain in one launch.
//
// Replaces the double loop of extract_iv_surface (engine/surface.py:69-126) and the per-strike loop of /api/smile
// (engine/app.py:226-234), each iteration of which runs SciPy's brentq on a Python objective (engine/surface.py:48-66:
// ~10 objective evaluations, each two scipy.stats.norm.cdf calls).  One thread per option, fp64:
//   f(sigma) = bs_call_price / bs_put_price (engine/surface.py:22-37, including the T <= 1e-10 or sigma <= 1e-10
//   intrinsic branch) - price;   no root in [lo, hi] (f(lo) f(hi) > 0, or a NaN anywhere) -> NaN (the reference's None);
//   otherwise a bracketed Newton iteration on the monotone f (vega > 0): Newton step when it stays inside the bracket,
//   bisection otherwise, until the bracket or the step is below 1e-13.  brentq's answer (xtol = 1e-8) lies within
//   1e-8 of the root; this one within ~1e-13.
#include "common.cuh"

namespace b200mc {

struct IvArgs {
    double S, r, q, lo, hi;
    int64_t n;
};

__device__ __forceinline__ double iv_price(double S, double K, double T, double r, double q, double sigma, bool call,
                                           double *vega)
{
    const double df_q = exp(-q * T), df_r = exp(-r * T);
    if (T <= 1e-10 || sigma <= 1e-10) {                                 // surface.py:24-25, 32-33
        if (vega) *vega = 0.0;
        // (separately rounded products, as NumPy evaluates them: an FMA here would leave a 1e-12 residue where the
        // reference finds the price EQUAL to the intrinsic value and returns lo)
        const double fwd = __dmul_rn(S, df_q), pvk = __dmul_rn(K, df_r);
        return call ? fmax(__dadd_rn(fwd, -pvk), 0.0) : fmax(__dadd_rn(pvk, -fwd), 0.0);
    }
    const double sT = sigma * sqrt(T);
    const double d1 = (log(S / K) + (r - q + 0.5 * sigma * sigma) * T) / sT;   // :26
    const double d2 = d1 - sT;
    if (vega) *vega = S * df_q * sqrt(T) * 0.3989422804014327 * exp(-0.5 * d1 * d1);   // :45
    return call ? __dadd_rn(__dmul_rn(__dmul_rn(S, df_q), normcdf(d1)), -__dmul_rn(__dmul_rn(K, df_r), normcdf(d2)))      // :28
                : __dadd_rn(__dmul_rn(__dmul_rn(K, df_r), normcdf(-d2)), -__dmul_rn(__dmul_rn(S, df_q), normcdf(-d1)));   // :37
}

__global__ void __launch_bounds__(128)
k_implied_vol(const __grid_constant__ IvArgs a, const double *__restrict__ price, const double *__restrict__ strike,
              const double *__restrict__ maturity, const int32_t *__restrict__ is_call, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double P = price[i], K = strike[i], T = maturity[i];
    const bool call = is_call[i] != 0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double lo = a.lo, hi = a.hi;
    double flo = iv_price(a.S, K, T, a.r, a.q, lo, call, nullptr) - P;
    double fhi = iv_price(a.S, K, T, a.r, a.q, hi, call, nullptr) - P;
    if (!(flo * fhi <= 0.0)) { out[i] = nan; return; }                  // surface.py:62-63 (NaN: brentq raises -> None)
    if (flo == 0.0) { out[i] = lo; return; }
    if (fhi == 0.0) { out[i] = hi; return; }
    const bool rising = fhi > 0.0;                                       // always, for a Black-Scholes price; kept general
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        double vega;
        const double fx = iv_price(a.S, K, T, a.r, a.q, x, call, &vega) - P;
        if (fx == 0.0) break;
        if ((fx > 0.0) == rising) hi = x; else lo = x;
        double xn = (vega > 0.0) ? x - fx / vega : nan;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);                 // Newton left the bracket (or no slope): bisect
        const double step = fabs(xn - x);
        x = xn;
        if (step < 1e-13 || hi - lo < 1e-13) break;
    }
    out[i] = x;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_implied_vol(b200mc_handle *h, int64_t n, const double *prices, const double *strikes,
                                  const double *maturities, const int32_t *is_call, double S, double r, double q,
                                  double lo, double hi, double *out_iv)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (n <= 0) return fail(h, B200MC_EINVAL, "n must be positive");
    if (!prices || !strikes || !maturities || !is_call || !out_iv) return fail(h, B200MC_EINVAL, "NULL array argument");
    if (!(lo < hi)) return fail(h, B200MC_EINVAL, "lo must be below hi");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    // staging: [price n][strike n][maturity n][iv n][is_call n (int32)]
    const size_t N = (size_t)n;
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, N * 8 * 4 + N * 4));
    double *dP = (double *)h->d_stage, *dK = dP + N, *dT = dK + N, *dO = dT + N;
    int32_t *dC = (int32_t *)(dO + N);
    B200MC_CUDA(h, cudaMemcpyAsync(dP, prices, N * 8, cudaMemcpyHostToDevice, h->stream));
    B200MC_CUDA(h, cudaMemcpyAsync(dK, strikes, N * 8, cudaMemcpyHostToDevice, h->stream));
    B200MC_CUDA(h, cudaMemcpyAsync(dT, maturities, N * 8, cudaMemcpyHostToDevice, h->stream));
    B200MC_CUDA(h, cudaMemcpyAsync(dC, is_call, N * 4, cudaMemcpyHostToDevice, h->stream));
    IvArgs a;
    a.S = S; a.r = r; a.q = q; a.lo = lo; a.hi = hi; a.n = n;
    k_implied_vol<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a, dP, dK, dT, dC, dO);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out_iv, dO, N * 8, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
