// risk.cu -- tail metrics of a P&L vector on the device.
//
// Restates compute_risk_metrics (engine/risk.py:117-155) and _hill_estimator (:158-173) without the full sort:
//   sorted[cutoff]            -> exact order statistic by most-significant-digit radix select over the order-preserving
//                                64-bit image of the doubles (8 passes of 8 bits; the vector stays L2 resident)
//   mean(sorted[:cutoff])     -> sum of the elements below the threshold + (cutoff - count_below) copies of it (ties)
//   Hill sum over the k largest losses -> same construction with the k-th order statistic
//   mean / std / skew / kurt  -> two passes (mean first, central moments second) like np.mean / np.std (:137-144)
// Index conventions are the reference's: cutoff = int(n (1 - confidence)) (:128), var = -sorted[cutoff] (:129),
// cvar = -mean(sorted[:cutoff]) (:130), k = max(int(sqrt(m)), 10) clipped to m - 1 over the m strictly negative
// returns, tail index only when m > 20 (:147-150,162-173).
#include <stdlib.h>

#include "common.cuh"

namespace b200mc {

constexpr int RK_THREADS = 256;

struct SelectState {
    unsigned long long prefix[2];
    long long rank[2];
    unsigned long long hist[2][256];
    double thr[2];
};

__device__ __forceinline__ unsigned long long key_of(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

template <typename T>
__global__ void __launch_bounds__(RK_THREADS)
k_risk_pass1(const T *__restrict__ x, int64_t n, unsigned long long *__restrict__ keys, double *partials,
             unsigned int *counter, double *out /* [2]: sum, count of x < 0 */)
{
    __shared__ double smem[(RK_THREADS / 32) * 2];
    double v[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RK_THREADS) {
        const double d = (double)x[i];
        keys[i] = key_of(d);
        v[0] += d;
        if (d < 0.0) v[1] += 1.0;
    }
    block_finish<2>(v, smem, partials, counter, out);
}

__global__ void __launch_bounds__(RK_THREADS)
k_risk_hist(const unsigned long long *__restrict__ keys, int64_t n, int pass, int nsel, SelectState *st)
{
    __shared__ unsigned int h[2][256];
    for (int i = threadIdx.x; i < 512; i += RK_THREADS) (&h[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 8 * pass;
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    for (int64_t i = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RK_THREADS) {
        const unsigned long long k = keys[i];
        const unsigned long long hi = (pass == 7) ? 0ull : (k >> (shift + 8));
        const unsigned int digit = (unsigned int)(k >> shift) & 255u;
        if (pass == 7 || hi == (p0 >> (shift + 8))) atomicAdd(&h[0][digit], 1u);
        if (nsel > 1 && (pass == 7 || hi == (p1 >> (shift + 8)))) atomicAdd(&h[1][digit], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256 * nsel; i += RK_THREADS) {
        const unsigned int c = (&h[0][0])[i];
        if (c) atomicAdd(&st->hist[0][0] + i, (unsigned long long)c);
    }
}

// One CTA of 256 threads, one bin each: block scan of the (global) histogram, the thread whose bin holds order statistic
// `rank` publishes it.  (A single thread walking the 256 bins took 10-24 us per digit: 8 digits = half of a sharded call.)
__global__ void __launch_bounds__(256) k_risk_pick(int pass, int nsel, SelectState *st)
{
    __shared__ unsigned long long wsum[8];
    __shared__ int s_bin;
    __shared__ unsigned long long s_below;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int s = 0; s < nsel; ++s) {
        const unsigned long long c = st->hist[s][t];
        const long long r = st->rank[s];
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += nb;
        }
        if (lane == 31) wsum[w] = inc;
        if (t == 0) s_bin = -1;
        __syncthreads();
        unsigned long long base = 0, total = 0;
        for (int i = 0; i < 8; ++i) {
            if (i < w) base += wsum[i];
            total += wsum[i];
        }
        inc += base;
        const unsigned long long exc = inc - c;
        if ((long long)exc <= r && (long long)inc > r) { s_bin = t; s_below = exc; }      // at most one thread
        __syncthreads();
        st->hist[s][t] = 0ull;
        if (t == 0) {
            const int bin = s_bin < 0 ? 255 : s_bin;                 // rank beyond the total: as the sequential walk ended
            const unsigned long long below = s_bin < 0 ? total : s_below;
            st->rank[s] = r - (long long)below;
            st->prefix[s] |= (unsigned long long)bin << (8 * pass);
            if (pass == 0) st->thr[s] = value_of(st->prefix[s]);
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(RK_THREADS)
k_risk_pass2(const T *__restrict__ x, int64_t n, double mean, int nsel, const SelectState *st, double *partials,
             unsigned int *counter, double *out /* [6] */)
{
    __shared__ double smem[(RK_THREADS / 32) * 6];
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double tc = st->thr[0], tk = nsel > 1 ? st->thr[1] : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RK_THREADS) {
        const double d = (double)x[i];
        const double c = d - mean, c2 = c * c;
        v[0] += c2;
        v[1] += c2 * c;
        v[2] += c2 * c2;
        if (d < tc) { v[3] += 1.0; v[4] += d; }
        if (nsel > 1 && d < tk) v[5] += log(d / tk);       // both negative: loss_i / loss_k   (:170)
    }
    block_finish<6>(v, smem, partials, counter, out);
}


// ---------------------------------------------------------------------------------------------- fused: ONE launch
// The same exact selection as the kernels above, but as one persistent cooperative kernel: no keys array, no host
// decision in the middle, the vector is read ONCE PER DIGIT (6 x 11-bit digits for float64 input, 3 for float32) instead
// of 10 times, and everything else rides on those passes:
//   pass 0   sum, count of negatives, histogram of the top digit (shared by both selections)
//   pass d   candidates of each selection (prefix equal so far) -> histogram of digit d; elements that fell BELOW the
//            prefix at the previous digit are, exactly once, added to that selection's tail sums (count, sum for CVaR,
//            sum of log|x| for Hill); pass 1 also accumulates the central moments (the mean is known after pass 0)
//   end      candidates that differ from the threshold only in the last digit all equal value(prefix | b): their tail
//            contribution comes from the last histogram alone
// Between passes: a grid barrier, then EVERY CTA picks the digit from the global histogram itself (2048 bins, one block
// scan) -- no second barrier, no single-thread walk (the one-thread k_risk_pick of round 1 was 10-24 us per pass: 150 of the 283 us a 4M-value
// call took).  Sums are folded in a fixed order (per-CTA partials, then one ordered sum), so a launch geometry is
// bitwise reproducible.  Measured on B200: see DESIGN.md section 4.4.
#ifndef B200MC_RISK_DEFAULT_CTAS
#define B200MC_RISK_DEFAULT_CTAS 2
#endif
constexpr int RF_THREADS = 512;
constexpr int RF_BITS = 11;
constexpr int RF_BINS = 1 << RF_BITS;
constexpr int RF_MAXPASS = 6;
constexpr int RF_NACC = 9;       // c2, c3, c4, cnt0, sum0, cnt1, logsum1, sum, neg

constexpr int RF_CAP = 2048;     // a selection with at most this many candidates left is finished by gathering them

struct FusedState {
    unsigned int hist[RF_MAXPASS][2][RF_BINS];
    unsigned long long barrier;  // grid barrier counter; the last CTA to leave the kernel resets it (and ncand, done)
    unsigned int ncand[2];
    unsigned int done;
    unsigned int pad_[11];
    double result[16];           // sum, neg, c2, c3, c4, cnt0, sum0, cnt1, logsum1, thr0, thr1
    unsigned long long cand[2][RF_CAP];
    unsigned long long trace[16];    // %globaltimer of CTA 0 at the phase boundaries (B200MC_RISK_TRACE=1 prints them)
};

template <typename T> struct RfKey;
template <> struct RfKey<double> {
    using K = unsigned long long;
    static constexpr int BITS = 64, NPASS = 6;
    __device__ static __forceinline__ K of(double x) { return key_of(x); }
    __device__ static __forceinline__ double value(K k) { return value_of(k); }
};
template <> struct RfKey<float> {
    using K = unsigned int;
    static constexpr int BITS = 32, NPASS = 3;
    __device__ static __forceinline__ K of(float x)
    {
        const unsigned int b = __float_as_uint(x);
        return (b >> 31) ? ~b : (b | 0x80000000u);
    }
    __device__ static __forceinline__ double value(K k)
    {
        const unsigned int b = (k >> 31) ? (k & 0x7fffffffu) : ~k;
        return (double)__uint_as_float(b);
    }
};
// digit d (0 = most significant) occupies bits [shift(d), shift(d) + width(d)): 11 bits each, the last one the remainder
template <typename T> __device__ __forceinline__ int rf_shift(int d)
{
    const int s = RfKey<T>::BITS - RF_BITS * (d + 1);
    return s > 0 ? s : 0;
}
template <typename T> __device__ __forceinline__ int rf_width(int d)
{
    const int hi = RfKey<T>::BITS - RF_BITS * d;
    return hi < RF_BITS ? hi : RF_BITS;
}

__device__ __forceinline__ void rf_grid_sync(unsigned long long *bar, unsigned long long &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1ull);
        while (*(volatile unsigned long long *)bar < target) { }
        __threadfence();
    }
    __syncthreads();
}

// One vote per DISTINCT bin and warp: lanes that hit the same bin (ties; the top digit of doubles, which is little more
// than sign + exponent) are counted with one shared-memory atomic instead of a serialised chain of them.
__device__ __forceinline__ void rf_vote(unsigned int *h, unsigned int digit)
{
    const unsigned int peers = __match_any_sync(__activemask(), digit);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[digit], (unsigned int)__popc(peers));
}

// grid-stride loop over x[0, n) with four independent 16-byte loads in flight per thread (the passes are latency bound
// otherwise: one 8-byte load per thread and iteration keeps ~1 MB in flight on the whole device, four keep 3.5-4 TB/s,
// four 16-byte ones 64 B per thread = 9.7 MB; requesting the next four before the current ones are consumed was measured
// and changes nothing, profiles/r02_risk_vec_probe.txt).  Elements before the first 16-byte boundary and after the last full
// vector are visited one by one.  The visiting ORDER of a thread is fixed by (n, alignment, grid) only, so the fp64 tail
// sums stay reproducible run to run.
template <typename T> struct RfVec;
template <> struct RfVec<double> { using V = double2; static constexpr int N = 2; };
template <> struct RfVec<float> { using V = float4; static constexpr int N = 4; };
template <typename F> __device__ __forceinline__ void rf_each(const double2 &v, F &body) { body(v.x); body(v.y); }
template <typename F> __device__ __forceinline__ void rf_each(const float4 &v, F &body) { body(v.x); body(v.y); body(v.z); body(v.w); }

template <typename T, typename F>
__device__ __forceinline__ void rf_foreach(const T *__restrict__ x, long long n, long long i0, long long stride, F body)
{
    using V = typename RfVec<T>::V;
    constexpr int NV = RfVec<T>::N;
    long long head = (long long)(((16u - (unsigned)((uintptr_t)x & 15u)) & 15u) / sizeof(T));
    if (head > n) head = n;
    const long long nvec = (n - head) / NV;
    const V *__restrict__ xv = reinterpret_cast<const V *>(x + head);
    long long i = i0;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const V a = xv[i], b = xv[i + stride], c = xv[i + 2 * stride], d = xv[i + 3 * stride];
        rf_each(a, body); rf_each(b, body); rf_each(c, body); rf_each(d, body);
    }
    for (; i < nvec; i += stride) { const V a = xv[i]; rf_each(a, body); }
    for (long long j = i0; j < head; j += stride) body(x[j]);
    for (long long j = head + nvec * NV + i0; j < n; j += stride) body(x[j]);
}

// deterministic sum of v over the CTA's threads -> every thread gets it
__device__ __forceinline__ double rf_block_sum(double v, double *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < RF_THREADS / 32; ++w) t += red[w];
    return t;
}

// exact sum of 64-bit integers over the CTA's threads (associative: the order of the additions cannot matter)
__device__ __forceinline__ long long rf_block_sum_ll(long long v, unsigned long long *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) red[warp] = (unsigned long long)v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < RF_THREADS / 32; ++w) t += (long long)red[w];
    return t;
}

// every CTA: find, in the global histogram h (nbins bins), the bin that holds order statistic `rank`; returns the bin and
// the number of elements in lower bins.  4 bins per thread + one block scan.
template <bool GLOBAL = true>
__device__ __forceinline__ void rf_pick(const unsigned int *h, int nbins, long long rank, unsigned long long *scan,
                                        int *bin_out, long long *below_out, long long *count_out)
{
    const int t = threadIdx.x;
    unsigned int c[4];
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = 4 * t + i;
        c[i] = b < nbins ? (GLOBAL ? __ldcg(h + b) : h[b]) : 0u;
        mine += c[i];
    }
    // inclusive scan of `mine` over the 512 threads
    const int lane = t & 31, warp = t >> 5;
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
    }
    __syncthreads();
    if (lane == 31) scan[warp] = inc;
    __syncthreads();
    unsigned long long before = 0;
    for (int w = 0; w < warp; ++w) before += scan[w];
    unsigned long long cum = before + inc - mine;                 // elements in bins below 4 t
    if (t == 0) { *bin_out = -1; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (c[i] != 0u && (long long)cum <= rank && rank < (long long)(cum + c[i])) { *bin_out = 4 * t + i; *below_out = (long long)cum; *count_out = (long long)c[i]; }
        cum += c[i];
    }
    __syncthreads();
    if (*bin_out < 0) {                                            // rank beyond the population (cannot happen for valid ranks)
        if (t == 0) { *bin_out = nbins - 1; *below_out = 0; *count_out = 0x7fffffffffffffffll; }
        __syncthreads();
    }
}

template <typename T, int MINB>
__global__ void __launch_bounds__(RF_THREADS, MINB)
k_risk_fused(const T *__restrict__ x, long long n, double confidence, FusedState *st, double *partials /* [grid][RF_NACC] */,
             unsigned long long bar_base)
{
    using KT = RfKey<T>;
    using K = typename KT::K;
    constexpr int NPASS = KT::NPASS;
    __shared__ unsigned int sh[2][RF_BINS];
    __shared__ double red[RF_THREADS / 32];
    __shared__ unsigned long long scan[RF_THREADS / 32];
    __shared__ int s_bin[2];
    __shared__ long long s_below[2], s_cnt[2];
    __shared__ double s_mean;
    __shared__ long long s_rank[2];
    __shared__ int s_nsel;
    const int tid = threadIdx.x;
    int trace_n = 0;
    auto mark = [&]() {
        if (blockIdx.x == 0 && tid == 0 && trace_n < 16) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            st->trace[trace_n++] = t;
        }
    };
    mark();
    unsigned long long bar_target = bar_base;
    const long long stride = (long long)gridDim.x * RF_THREADS, i0 = (long long)blockIdx.x * RF_THREADS + tid;

    // ---- pass 0 -------------------------------------------------------------------------------------------------
    for (int i = tid; i < RF_BINS; i += RF_THREADS) sh[0][i] = 0u;
    __syncthreads();
    double sum = 0.0, neg = 0.0;
    {
        const int shift = rf_shift<T>(0);
        rf_foreach<T>(x, n, i0, stride, [&](const T xv) {
            const double d = (double)xv;
            sum += d;
            if (d < 0.0) neg += 1.0;
            rf_vote(&sh[0][0], (unsigned int)(KT::of(xv) >> shift));
        });
    }
    __syncthreads();
    mark();                                                                      // 1: pass 0 read
    for (int i = tid; i < RF_BINS; i += RF_THREADS)
        if (sh[0][i]) atomicAdd(&st->hist[0][0][i], sh[0][i]);
    {
        const double bs = rf_block_sum(sum, red), bn = rf_block_sum(neg, red);
        if (tid == 0) { partials[(size_t)blockIdx.x * RF_NACC + 7] = bs; partials[(size_t)blockIdx.x * RF_NACC + 8] = bn; }
    }
    rf_grid_sync(&st->barrier, bar_target);
    mark();                                                                      // 2: first grid barrier passed
    // every CTA: total sum and count of negatives in a fixed order, then the two ranks (engine/risk.py:128-134,147-166)
    {
        double ts = 0.0, tn = 0.0;
        for (unsigned int b = tid; b < gridDim.x; b += RF_THREADS) {
            ts += __ldcg(&partials[(size_t)b * RF_NACC + 7]);
            tn += __ldcg(&partials[(size_t)b * RF_NACC + 8]);
        }
        ts = rf_block_sum(ts, red);
        tn = rf_block_sum(tn, red);
        if (tid == 0) {
            s_mean = ts / (double)n;                                             // :137
            const long long m = (long long)tn;                                   // len(losses), :147
            long long cutoff = (long long)((double)n * (1.0 - confidence));      // :128
            if (cutoff < 0) cutoff = 0;
            s_rank[0] = cutoff < n ? cutoff : 0;                                 // :129
            long long k = 0;
            const bool want_hill = m > 20;                                       // :150
            if (want_hill) {
                k = (long long)sqrt((double)m);                                  // :165
                if (k < 10) k = 10;
                if (k > m - 1) k = m - 1;                                        // :166
            }
            s_rank[1] = k;
            s_nsel = want_hill ? 2 : 1;
            if (blockIdx.x == 0) { st->result[0] = ts; st->result[1] = tn; }
        }
        __syncthreads();
    }
    const double mean = s_mean;
    const int nsel = s_nsel;
    K prefix[2] = {0, 0};            // decided digits of each selection, in place (lower bits zero)
    long long rank[2] = {s_rank[0], s_rank[1]};
    // gather: once every selection has at most RF_CAP candidates left, the next pass collects them instead of counting
    // a digit and CTA 0 finishes the selections with a sort in shared memory -- 3 reads of the vector for 4e6 continuous
    // values instead of 6.  Heavy ties (an option P&L vector is 45 % one value) never get there and take all the digits.
    bool gather = true;
    for (int sel = 0; sel < nsel; ++sel) {
        rf_pick(&st->hist[0][0][0], 1 << rf_width<T>(0), rank[sel], scan, &s_bin[sel], &s_below[sel], &s_cnt[sel]);
        prefix[sel] = (K)s_bin[sel] << rf_shift<T>(0);
        rank[sel] -= s_below[sel];
        gather = gather && s_cnt[sel] <= RF_CAP;
        __syncthreads();
    }
    bool finished = false;
    mark();                                                                      // 3: ranks + digit-0 pick

    // ---- passes 1 .. NPASS - 1 ---------------------------------------------------------------------------------
    double c2 = 0.0, c3 = 0.0, c4 = 0.0, cnt[2] = {0.0, 0.0}, acc[2] = {0.0, 0.0};     // acc: sum (sel 0), sum log|x| (sel 1)
#pragma unroll 1
    for (int d = 1; d < NPASS; ++d) {
        const bool gpass = gather;                                               // the same on every CTA
        for (int i = tid; i < 2 * RF_BINS; i += RF_THREADS) (&sh[0][0])[i] = 0u;
        __syncthreads();
        const int shift = rf_shift<T>(d), up = rf_shift<T>(d - 1);              // up: everything decided so far sits above it
        const unsigned int mask = (1u << rf_width<T>(d)) - 1u;
        const K p0 = prefix[0] >> up, p1 = prefix[1] >> up;
        // decided before the previous digit (d >= 2): equality there makes an element "newly below" at digit d - 1
        const int up2 = d >= 2 ? rf_shift<T>(d - 2) : 0;
        const K q0 = d >= 2 ? (prefix[0] >> up2) : 0, q1 = d >= 2 ? (prefix[1] >> up2) : 0;
        double cnt0 = 0.0, acc0 = 0.0, cnt1 = 0.0, acc1 = 0.0;                   // this pass's tail sums, in registers
        rf_foreach<T>(x, n, i0, stride, [&](const T xv) {
            const double v = (double)xv;
            const K key = KT::of(xv);
            if (d == 1) {
                const double c = v - mean, cc = c * c;
                c2 += cc; c3 += cc * c; c4 += cc * cc;
            }
            const K hi = key >> up;
            const unsigned int digit = (unsigned int)(key >> shift) & mask;
            if (hi == p0) {
                if (gpass) st->cand[0][atomicAdd(&st->ncand[0], 1u)] = (unsigned long long)key;
                else rf_vote(&sh[0][0], digit);
            } else if (hi < p0 && (d == 1 || (key >> up2) == q0)) { cnt0 += 1.0; acc0 += v; }
            if (nsel > 1) {
                if (hi == p1) {
                    if (gpass) st->cand[1][atomicAdd(&st->ncand[1], 1u)] = (unsigned long long)key;
                    else rf_vote(&sh[1][0], digit);
                } else if (hi < p1 && (d == 1 || (key >> up2) == q1)) { cnt1 += 1.0; acc1 += log(fabs(v)); }
            }
        });
        cnt[0] += cnt0; acc[0] += acc0; cnt[1] += cnt1; acc[1] += acc1;
        __syncthreads();
        mark();                                                                  // pass d read
        if (!gpass) {
            for (int i = tid; i < nsel * RF_BINS; i += RF_THREADS) {
                const unsigned int c = (&sh[0][0])[i];
                if (c) atomicAdd(&st->hist[d][0][0] + i, c);
            }
        }
        rf_grid_sync(&st->barrier, bar_target);
        mark();                                                                  // barrier after pass d
        if (gpass) {
            // CTA 0 finishes each selection on its <= RF_CAP candidates alone: the remaining digits are counted in shared
            // memory (at most 4 keys per thread), a microsecond per digit.  (A bitonic sort of the 2048 keys in
            // shared memory took 23 us per selection.)
            if (blockIdx.x == 0) {
                K cand_lo[2], cand_hi[2];                                    // the key range the gathered candidates lie in
                for (int sel = 0; sel < nsel; ++sel) {
                    cand_lo[sel] = prefix[sel];
                    cand_hi[sel] = prefix[sel] | (((K)1 << up) - (K)1);
                }
                for (int sel = 0; sel < nsel; ++sel) {
                    const int nc = (int)__ldcg(&st->ncand[sel]);
                    for (int dd = d; dd < NPASS; ++dd) {
                        __syncthreads();
                        for (int i = tid; i < RF_BINS; i += RF_THREADS) sh[0][i] = 0u;
                        __syncthreads();
                        const int sh_d = rf_shift<T>(dd), up_d = rf_shift<T>(dd - 1);
                        const unsigned int mask_d = (1u << rf_width<T>(dd)) - 1u;
                        for (int i = tid; i < nc; i += RF_THREADS) {           // the keys are re-read from L2: <= 4 per thread
                            const K key = (K)__ldcg(&st->cand[sel][i]);
                            if ((key >> up_d) == (prefix[sel] >> up_d))
                                atomicAdd(&sh[0][(unsigned int)(key >> sh_d) & mask_d], 1u);
                        }
                        __syncthreads();
                        rf_pick<false>(&sh[0][0], 1 << rf_width<T>(dd), rank[sel], scan, &s_bin[sel], &s_below[sel], &s_cnt[sel]);
                        prefix[sel] |= (K)s_bin[sel] << sh_d;
                        rank[sel] -= s_below[sel];
                    }
                    const K thr_key = prefix[sel];
                    // The candidates arrive in the order of a global atomic counter, so their tail sum must not depend on
                    // the order of its additions: every term is turned into a 64-bit fixed-point integer, summed exactly
                    // (high and low halves apart: 2048 terms cannot overflow either) and rounded ONCE.  Values: the
                    // candidates share their leading digit(s), i.e. sign and all but the last exponent bit, so with the
                    // scale taken from the largest magnitude the prefix allows v 2^sc is an exact integer below 2^62.
                    // Logarithms (|log| < 2^10): scale 2^50, exact from |log| >= 4 and within 2^-51 below.
                    int sc = 50;
                    bool exact = true;
                    if (sel == 0) {
                        const double vmax = fmax(fabs(KT::value(cand_lo[sel])), fabs(KT::value(cand_hi[sel])));
                        exact = vmax < 1.7976931348623157e308;            // (an infinity among the inputs: plain sums)
                        sc = (exact && vmax > 0.0) ? 61 - ilogb(vmax) : 0;
                    }
                    double cb = 0.0, ab = 0.0;
                    long long shi = 0, slo = 0;
                    for (int i = tid; i < nc; i += RF_THREADS) {
                        const K key = (K)__ldcg(&st->cand[sel][i]);
                        if (key < thr_key) {
                            const double v = KT::value(key);
                            const double x = sel == 0 ? v : log(fabs(v));
                            cb += 1.0;
                            if (exact) {
                                const long long q = llrint(scalbn(x, sc));
                                slo += q & 0xffffffffll;
                                shi += q >> 32;
                            } else {
                                ab += x;
                            }
                        }
                    }
                    cnt[sel] += cb;
                    if (exact) {
                        const long long thi = rf_block_sum_ll(shi, scan), tlo = rf_block_sum_ll(slo, scan);
                        if (tid == 0) acc[sel] += scalbn((double)thi * 4294967296.0 + (double)tlo, -sc);
                    } else {
                        acc[sel] += ab;
                    }
                }
                __syncthreads();
            }
            mark();                                                              // candidates resolved
            finished = true;
            break;
        }
        gather = true;
        for (int sel = 0; sel < nsel; ++sel) {
            rf_pick(&st->hist[d][sel][0], 1 << rf_width<T>(d), rank[sel], scan, &s_bin[sel], &s_below[sel], &s_cnt[sel]);
            prefix[sel] |= (K)s_bin[sel] << shift;
            rank[sel] -= s_below[sel];
            gather = gather && s_cnt[sel] <= RF_CAP;
            __syncthreads();
        }
        mark();                                                                  // digit d picked
    }
    // ---- the candidates that differ from the threshold in the last digit only: from the last histogram (CTA 0) ------
    if (blockIdx.x == 0 && !finished) {
        constexpr int d = NPASS - 1;
        const int nb = 1 << rf_width<T>(d);
        for (int sel = 0; sel < nsel; ++sel) {
            const int chosen = (int)(prefix[sel] & (K)(nb - 1));
            const K base = prefix[sel] & ~(K)(nb - 1);
            for (int b = tid; b < chosen; b += RF_THREADS) {
                const unsigned int c = __ldcg(&st->hist[d][sel][b]);
                if (c) {
                    const double v = KT::value(base | (K)b);
                    cnt[sel] += (double)c;
                    acc[sel] += (double)c * (sel == 0 ? v : log(fabs(v)));
                }
            }
        }
    }
    // ---- fold: per-CTA partials in a fixed order, then CTA 0 sums them -------------------------------------------
    {
        const double v7[7] = {c2, c3, c4, cnt[0], acc[0], cnt[1], acc[1]};
        for (int j = 0; j < 7; ++j) {
            const double b = rf_block_sum(v7[j], red);
            if (tid == 0) partials[(size_t)blockIdx.x * RF_NACC + j] = b;
        }
    }
    rf_grid_sync(&st->barrier, bar_target);
    mark();                                                                      // final barrier passed
    // leave the histograms zeroed for the next launch (all CTAs are past their last read of them)
    for (long long i = i0; i < (long long)(RF_MAXPASS * 2 * RF_BINS); i += stride) (&st->hist[0][0][0])[i] = 0u;
    if (blockIdx.x == 0) {
        for (int j = 0; j < 7; ++j) {
            double t = 0.0;
            for (unsigned int b = tid; b < gridDim.x; b += RF_THREADS) t += __ldcg(&partials[(size_t)b * RF_NACC + j]);
            t = rf_block_sum(t, red);
            if (tid == 0) st->result[2 + j] = t;
        }
        if (tid == 0) {
            st->result[9] = KT::value(prefix[0]);
            st->result[10] = nsel > 1 ? KT::value(prefix[1]) : 0.0;
            st->result[11] = (double)nsel;
            st->result[12] = finished ? 1.0 : 0.0;
        }
    }
    mark();                                                                      // end of CTA 0
    if (blockIdx.x == 0 && tid == 0) { for (int i = trace_n; i < 16; ++i) st->trace[i] = 0ull; }
    // the last CTA to get here re-arms the barrier and the candidate counters (nobody can still be spinning: every CTA
    // has left the final barrier before it signs off)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&st->done, 1u) == gridDim.x - 1) {
            st->barrier = 0ull;
            st->ncand[0] = st->ncand[1] = 0u;
            st->done = 0u;
        }
    }
}

__global__ void k_set_double(double *p, double v) { *p = v; }

// sharded = the vector is the LOCAL shard of a vector spread over the ranks of the handle's peer connection (peer.cu):
// the two sums of pass 1 (+ the shard length), the histograms of every radix pass and the six sums of pass 2 are
// all-reduced on the device between the kernels, so every rank takes the same decisions and returns the GLOBAL metrics;
// the shards never move and the host synchronises twice, as in the single-device case.  n may be 0 on a rank.
// Single-device case: ONE cooperative launch + one read-back (k_risk_fused).  Returns 1 when the fused path cannot run
// (no cooperative launch, more than 2^32 - 1 elements) so that the caller takes the multi-kernel path.
template <typename T>
static int risk_run_fused(b200mc_handle *h, const T *x_dev, const int64_t n, double confidence, double out[8])
{
    if (n >= (int64_t)4294967295ll) return 1;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!coop) return 1;
    // 1 CTA per SM with ~110 registers (no spills) or 2 with 64 (more loads in flight): tuning knob B200MC_RISK_CTAS
    const char *e = getenv("B200MC_RISK_CTAS");
    const bool two = e ? atoi(e) == 2 : (B200MC_RISK_DEFAULT_CTAS == 2);
    const void *kern = two ? (const void *)k_risk_fused<T, 2> : (const void *)k_risk_fused<T, 1>;
    int occ = 0;
    for (int i = 0; i < h->n_occ; ++i)
        if (h->occ_kern[i] == kern && h->occ_smem[i] == 0) occ = h->occ_val[i];
    if (occ == 0) {
        B200MC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RF_THREADS, 0));
        if (occ < 1) return 1;
        const int slot = h->n_occ < 64 ? h->n_occ++ : 63;
        h->occ_kern[slot] = kern;
        h->occ_smem[slot] = 0;
        h->occ_val[slot] = occ;
    }
    int64_t grid = (n + RF_THREADS - 1) / RF_THREADS;
    const int64_t cap = (int64_t)h->sm_count * (occ < 2 ? occ : 2);          // co-resident by construction
    if (grid > cap) grid = cap;
    const size_t off_pa = (sizeof(FusedState) + 255) & ~(size_t)255;
    // the select's state has its own buffer: the kernel leaves it zeroed when it ends, so only the first call (and a call
    // after a failed launch) pays for a memset
    if (!h->risk_state) {
        B200MC_CUDA(h, cudaMalloc(&h->risk_state, off_pa + (size_t)h->sm_count * 2 * RF_NACC * 8 + 64));
        h->risk_state_clean = false;
    }
    FusedState *st = (FusedState *)h->risk_state;
    double *partials = (double *)((char *)h->risk_state + off_pa);
    if (!h->risk_state_clean) B200MC_CUDA(h, cudaMemsetAsync(st, 0, sizeof(FusedState), h->stream));
    h->risk_state_clean = false;
    long long nn = n;
    unsigned long long bar_base = 0;                       // the kernel re-arms its barrier before it ends
    void *args[] = {(void *)&x_dev, (void *)&nn, (void *)&confidence, (void *)&st, (void *)&partials, (void *)&bar_base};
    B200MC_CUDA(h, cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(RF_THREADS), args, 0,
                                                h->stream));
    h->launches += 1;
    B200MC_TRY(ensure(h, &h->h_result, &h->h_result_bytes, 4096, true));
    double *r = (double *)h->h_result;                                        // pinned landing buffer
    B200MC_CUDA(h, cudaMemcpyAsync(r, st->result, 13 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->risk_state_clean = true;                                               // the kernel ran to its end
    if (getenv("B200MC_RISK_TRACE")) {
        unsigned long long tr[16];
        B200MC_CUDA(h, cudaMemcpy(tr, st->trace, sizeof(tr), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[k_risk_fused n=%lld grid=%lld] us since start:", (long long)n, (long long)grid);
        for (int i = 1; i < 16 && tr[i]; ++i) fprintf(stderr, " %.1f", (double)(tr[i] - tr[0]) * 1e-3);
        fprintf(stderr, "\n");
    }
    const double nnd = (double)n, mean = r[0] / nnd;                          // :137
    const int64_t m = (int64_t)r[1];
    int64_t cutoff = (int64_t)(nnd * (1.0 - confidence));                     // :128
    if (cutoff < 0) cutoff = 0;
    const bool want_hill = m > 20;                                            // :150
    int64_t k = 0;
    if (want_hill) {
        k = (int64_t)sqrt((double)m);                                         // :165
        if (k < 10) k = 10;
        if (k > m - 1) k = m - 1;                                             // :166
    }
    const double sd = sqrt(r[2] / nnd);                                       // :138 (np.std, ddof = 0)
    const double sdc = sd > 1e-10 ? sd : 1e-10;                               // :141
    const double skew = (r[3] / nnd) / (sdc * sdc * sdc);                     // :143
    const double kurt = (r[4] / nnd) / (sdc * sdc * sdc * sdc);               // :144
    const double thr0 = r[9], thr1 = r[10];
    double cvar;
    if (cutoff <= 0) cvar = -thr0;                                            // :130, else-branch (-sorted[0])
    else if (cutoff >= n) cvar = -mean;                                       // slice covers everything
    else cvar = -(r[6] + ((double)cutoff - r[5]) * thr0) / (double)cutoff;    // ties at the threshold
    double tail = NAN;
    if (want_hill && thr1 < 0.0) {
        const double logsum = r[8] - r[7] * log(fabs(thr1));                  // sum over x < thr1 of log(x / thr1)   (:170)
        if (logsum > 0.0 && r[7] > 0.0) tail = (double)k / logsum;            // :168-173
    }
    out[0] = -thr0; out[1] = cvar; out[2] = skew; out[3] = kurt; out[4] = kurt - 3.0; out[5] = tail;
    out[6] = mean; out[7] = sd;
    return 0;
}

template <typename T>
static int risk_run(b200mc_handle *h, const T *x_dev, const int64_t n_loc, double confidence, double out[8],
                    bool sharded = false)
{
    if (!sharded && n_loc > 0 && !getenv("B200MC_RISK_MULTIKERNEL")) {
        const int rc = risk_run_fused<T>(h, x_dev, n_loc, confidence, out);
        if (rc != 1) return rc;
    }
    int64_t n = n_loc;                                 // becomes the GLOBAL length after pass 1 when sharded
    int64_t grid = (n_loc + RK_THREADS - 1) / RK_THREADS;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    // scratch: [keys n*8][SelectState][partials grid*6*8][results 8*8]
    const size_t off_st = ((size_t)n * 8 + 255) & ~(size_t)255;
    const size_t off_pa = off_st + ((sizeof(SelectState) + 255) & ~(size_t)255);
    const size_t off_re = off_pa + (size_t)grid * 6 * 8;
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, off_re + 64));
    char *sc = (char *)h->d_scratch;
    unsigned long long *keys = (unsigned long long *)sc;
    SelectState *st = (SelectState *)(sc + off_st);
    double *partials = (double *)(sc + off_pa), *res = (double *)(sc + off_re);

    k_risk_pass1<T><<<(unsigned)grid, RK_THREADS, 0, h->stream>>>(x_dev, n_loc, keys, partials, h->d_counter, res);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    double r1[3] = {0.0, 0.0, (double)n_loc};
    if (sharded) {
        k_set_double<<<1, 1, 0, h->stream>>>(res + 2, (double)n_loc);
        B200MC_TRY(peer_allreduce_async(h, res, 3, false));
    }
    B200MC_CUDA(h, cudaMemcpyAsync(r1, res, sharded ? 24 : 16, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (sharded) B200MC_TRY(peer_check(h));
    n = (int64_t)r1[2];                                                      // from here on: the GLOBAL length
    if (n <= 0) return fail(h, B200MC_EINVAL, "returns must not be empty");
    const double mean = r1[0] / (double)n;                                   // :137
    const int64_t m = (int64_t)r1[1];                                        // len(losses), :147

    int64_t cutoff = (int64_t)((double)n * (1.0 - confidence));              // :128
    if (cutoff < 0) cutoff = 0;
    const int64_t rank_c = cutoff < n ? cutoff : 0;                          // :129
    int64_t k = 0;
    const bool want_hill = m > 20;                                           // :150
    if (want_hill) {
        k = (int64_t)sqrt((double)m);                                        // :165
        if (k < 10) k = 10;
        if (k > m - 1) k = m - 1;                                            // :166
    }
    const int nsel = want_hill ? 2 : 1;
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.rank[0] = rank_c;
    init.rank[1] = k;
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, sizeof(SelectState) > 4096 ? sizeof(SelectState) : 4096, true));
    memcpy(h->h_pinned, &init, sizeof(init));
    B200MC_CUDA(h, cudaMemcpyAsync(st, h->h_pinned, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    for (int pass = 7; pass >= 0; --pass) {
        k_risk_hist<<<(unsigned)grid, RK_THREADS, 0, h->stream>>>(keys, n_loc, pass, nsel, st);
        if (sharded) B200MC_TRY(peer_allreduce_async(h, &st->hist[0][0], 256 * nsel, true));
        k_risk_pick<<<1, 256, 0, h->stream>>>(pass, nsel, st);
        h->launches += 2;
    }
    B200MC_CUDA(h, cudaGetLastError());
    k_risk_pass2<T><<<(unsigned)grid, RK_THREADS, 0, h->stream>>>(x_dev, n_loc, mean, nsel, st, partials, h->d_counter, res);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (sharded) B200MC_TRY(peer_allreduce_async(h, res, 6, false));
    double r2[6], thr[2];
    B200MC_CUDA(h, cudaMemcpyAsync(r2, res, 48, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaMemcpyAsync(thr, &st->thr[0], 16, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (sharded) B200MC_TRY(peer_check(h));

    const double nn = (double)n;
    const double var_pop = r2[0] / nn;
    const double sd = sqrt(var_pop);                                         // :138 (np.std, ddof = 0)
    const double sdc = sd > 1e-10 ? sd : 1e-10;                              // :141
    const double skew = (r2[1] / nn) / (sdc * sdc * sdc);                    // :143
    const double kurt = (r2[2] / nn) / (sdc * sdc * sdc * sdc);              // :144
    const double var = -thr[0];                                              // :129
    double cvar;
    if (cutoff <= 0) cvar = -thr[0];                                         // :130, else-branch (-sorted[0])
    else if (cutoff >= n) cvar = -mean;                                      // slice covers everything
    else cvar = -(r2[4] + ((double)cutoff - r2[3]) * thr[0]) / (double)cutoff;
    double tail = NAN;
    if (want_hill && thr[1] < 0.0 && r2[5] > 0.0) tail = (double)k / r2[5];  // :168-173
    out[0] = var; out[1] = cvar; out[2] = skew; out[3] = kurt; out[4] = kurt - 3.0; out[5] = tail;
    out[6] = mean; out[7] = sd;
    return 0;
}

// Discounted option P&L from terminal spots, element-wise on the device: pnl[i] = discount * payoff(S[i * stride]) - premium
// (BASELINE config 4: "then a10 on D*max(S_T - K, 0) - premium").  stride > 1 reads a column of a path matrix in place.
template <typename TI, typename TO>
__global__ void __launch_bounds__(RK_THREADS)
k_option_pnl(const TI *__restrict__ S, int64_t n, int64_t stride, double strike, int is_call, double discount,
             double premium, TO *__restrict__ pnl)
{
    for (int64_t i = (int64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RK_THREADS) {
        const double s = (double)S[i * stride];
        const double pay = is_call ? fmax(s - strike, 0.0) : fmax(strike - s, 0.0);
        pnl[i] = (TO)(discount * pay - premium);
    }
}

// scratch layout shared by the single-call and the multi-rank entry points
struct RiskScratch {
    unsigned long long *keys;
    SelectState *st;
    double *partials, *res;
    int64_t grid;
};
static int risk_scratch(b200mc_handle *h, int64_t n, RiskScratch &r)
{
    r.grid = (n + RK_THREADS - 1) / RK_THREADS;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (r.grid > cap) r.grid = cap;
    const size_t off_st = ((size_t)n * 8 + 255) & ~(size_t)255;
    const size_t off_pa = off_st + ((sizeof(SelectState) + 255) & ~(size_t)255);
    const size_t off_re = off_pa + (size_t)r.grid * 6 * 8;
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, off_re + 64));
    char *sc = (char *)h->d_scratch;
    r.keys = (unsigned long long *)sc;
    r.st = (SelectState *)(sc + off_st);
    r.partials = (double *)(sc + off_pa);
    r.res = (double *)(sc + off_re);
    return 0;
}

} // namespace b200mc
using namespace b200mc;

extern "C" int b200mc_option_pnl(b200mc_handle *h, const void *S_dev, int64_t n, int64_t stride, int dtype_in, double strike,
                                 int is_call, double discount, double premium, int dtype_out, void *pnl_dev)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!S_dev || !pnl_dev || n <= 0 || stride <= 0) return fail(h, B200MC_EINVAL, "bad argument");
    if ((dtype_in != B200MC_F32 && dtype_in != B200MC_F64) || (dtype_out != B200MC_F32 && dtype_out != B200MC_F64))
        return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    int64_t grid = (n + RK_THREADS - 1) / RK_THREADS;
    const int64_t cap = (int64_t)h->sm_count * 16;
    if (grid > cap) grid = cap;
#define PNL(TI, TO) k_option_pnl<TI, TO><<<(unsigned)grid, RK_THREADS, 0, h->stream>>>((const TI *)S_dev, n, stride, strike, is_call, discount, premium, (TO *)pnl_dev)
    if (dtype_in == B200MC_F32) { if (dtype_out == B200MC_F32) PNL(float, float); else PNL(float, double); }
    else { if (dtype_out == B200MC_F32) PNL(double, float); else PNL(double, double); }
#undef PNL
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

// ---- multi-rank primitives: every rank holds a shard of the P&L vector; the host all-reduces the tiny results ----------
// (1) begin: order-preserving keys of the local shard + local { sum, count of negatives }.
extern "C" int b200mc_risk_begin(b200mc_handle *h, const void *pnl, int64_t n, int dtype, int on_device, double out[2])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || n < 0 || (n > 0 && !pnl)) return fail(h, B200MC_EINVAL, "bad argument");
    if (dtype != B200MC_F32 && dtype != B200MC_F64) return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    h->risk_x = nullptr; h->risk_n = n; h->risk_dtype = dtype;
    out[0] = out[1] = 0.0;
    if (n == 0) return 0;
    const size_t esz = dtype == B200MC_F64 ? 8 : 4;
    const void *x = pnl;
    if (!on_device) {
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n * esz + 256));
        B200MC_CUDA(h, cudaMemcpyAsync(h->d_stage, pnl, (size_t)n * esz, cudaMemcpyHostToDevice, h->stream));
        x = h->d_stage;
    }
    RiskScratch r;
    B200MC_TRY(risk_scratch(h, n, r));
    if (dtype == B200MC_F64)
        k_risk_pass1<double><<<(unsigned)r.grid, RK_THREADS, 0, h->stream>>>((const double *)x, n, r.keys, r.partials, h->d_counter, r.res);
    else
        k_risk_pass1<float><<<(unsigned)r.grid, RK_THREADS, 0, h->stream>>>((const float *)x, n, r.keys, r.partials, h->d_counter, r.res);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, r.res, 16, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->risk_x = x;
    return 0;
}

// (2) one radix pass: local 256-bin histograms of digit `pass` (7 = most significant byte) among the keys whose higher
// digits equal those of prefix[s], for nsel (1 or 2) concurrent selections.  hist = [2][256] counts.
extern "C" int b200mc_risk_hist(b200mc_handle *h, int pass, int nsel, const uint64_t prefix[2], uint64_t hist[512])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!prefix || !hist || pass < 0 || pass > 7 || nsel < 1 || nsel > 2) return fail(h, B200MC_EINVAL, "bad argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    memset(hist, 0, 512 * sizeof(uint64_t));
    if (h->risk_n == 0) return 0;
    if (!h->risk_x) return fail(h, B200MC_EINVAL, "b200mc_risk_begin has not been called");
    RiskScratch r;
    B200MC_TRY(risk_scratch(h, h->risk_n, r));
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.prefix[0] = prefix[0];
    init.prefix[1] = prefix[1];
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, sizeof(SelectState) > 4096 ? sizeof(SelectState) : 4096, true));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(h->h_pinned, &init, sizeof(init));
    B200MC_CUDA(h, cudaMemcpyAsync(r.st, h->h_pinned, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    k_risk_hist<<<(unsigned)r.grid, RK_THREADS, 0, h->stream>>>(r.keys, h->risk_n, pass, nsel, r.st);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(hist, &r.st->hist[0][0], 512 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// (3) finish: local { sum d^2, sum d^3, sum d^4, count(x < thr0), sum(x < thr0), sum log(x / thr1) over x < thr1 } with
// d = x - mean (the GLOBAL mean) and the GLOBAL thresholds.
extern "C" int b200mc_risk_finish(b200mc_handle *h, double mean, int nsel, const double thr[2], double out[6])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!thr || !out || nsel < 1 || nsel > 2) return fail(h, B200MC_EINVAL, "bad argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    for (int i = 0; i < 6; ++i) out[i] = 0.0;
    if (h->risk_n == 0) return 0;
    if (!h->risk_x) return fail(h, B200MC_EINVAL, "b200mc_risk_begin has not been called");
    RiskScratch r;
    B200MC_TRY(risk_scratch(h, h->risk_n, r));
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.thr[0] = thr[0];
    init.thr[1] = thr[1];
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, sizeof(SelectState) > 4096 ? sizeof(SelectState) : 4096, true));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(h->h_pinned, &init, sizeof(init));
    B200MC_CUDA(h, cudaMemcpyAsync(r.st, h->h_pinned, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    if (h->risk_dtype == B200MC_F64)
        k_risk_pass2<double><<<(unsigned)r.grid, RK_THREADS, 0, h->stream>>>((const double *)h->risk_x, h->risk_n, mean, nsel, r.st, r.partials, h->d_counter, r.res);
    else
        k_risk_pass2<float><<<(unsigned)r.grid, RK_THREADS, 0, h->stream>>>((const float *)h->risk_x, h->risk_n, mean, nsel, r.st, r.partials, h->d_counter, r.res);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, r.res, 48, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mc_risk_metrics(b200mc_handle *h, const void *pnl, int64_t n, int dtype, int on_device,
                                   double confidence, double out[8])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!pnl || !out) return fail(h, B200MC_EINVAL, "NULL argument");
    if (n <= 0) return fail(h, B200MC_EINVAL, "returns must not be empty");
    if (dtype != B200MC_F32 && dtype != B200MC_F64) return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    if (!(confidence == confidence)) return fail(h, B200MC_EINVAL, "confidence is NaN");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const size_t esz = dtype == B200MC_F64 ? 8 : 4;
    const void *x = pnl;
    if (!on_device) {
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n * esz + 256));
        B200MC_CUDA(h, cudaMemcpyAsync(h->d_stage, pnl, (size_t)n * esz, cudaMemcpyHostToDevice, h->stream));
        x = h->d_stage;
    }
    if (dtype == B200MC_F64) return risk_run<double>(h, (const double *)x, n, confidence, out);
    return risk_run<float>(h, (const float *)x, n, confidence, out);
}

extern "C" int b200mc_risk_metrics_sharded(b200mc_handle *h, const void *pnl_local, int64_t n_local, int dtype, int on_device,
                                           double confidence, double out[8])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || (n_local > 0 && !pnl_local)) return fail(h, B200MC_EINVAL, "NULL argument");
    if (n_local < 0) return fail(h, B200MC_EINVAL, "n_local must be >= 0");
    if (dtype != B200MC_F32 && dtype != B200MC_F64) return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    if (!(confidence == confidence)) return fail(h, B200MC_EINVAL, "confidence is NaN");
    if (h->peer_world < 1) return fail(h, B200MC_EINVAL, "call b200mc_peer_connect first");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const size_t esz = dtype == B200MC_F64 ? 8 : 4;
    const void *x = pnl_local;
    if (!on_device && n_local > 0) {
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n_local * esz + 256));
        B200MC_CUDA(h, cudaMemcpyAsync(h->d_stage, pnl_local, (size_t)n_local * esz, cudaMemcpyHostToDevice, h->stream));
        x = h->d_stage;
    }
    if (dtype == B200MC_F64) return risk_run<double>(h, (const double *)x, n_local, confidence, out, true);
    return risk_run<float>(h, (const float *)x, n_local, confidence, out, true);
}
