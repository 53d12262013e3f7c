// european.cu -- the fused European kernel: Philox draws in registers, log-Euler evolution, payoff and Greek
// sums reduced on chip.  Replaces, for the pseudo-random case, the RNG front end + both kernel runs + the NumPy
// reductions of MonteCarloEngine.price / price_batch (engine/monte_carlo.py:273-450) and the three re-simulations
// each of GreeksEngine.delta / vega / gamma (engine/greeks.py:53-203).  No path matrix touches HBM: the only
// global traffic is one 128-byte partial per (CTA, strike) and the final b200mc_sums; strikes travel in the kernel
// parameters, so a call makes no host->device copy and no stream synchronisation.
//
// CTA = 256 threads, one persistent wave.
//   one strike    every thread simulates its own paths (sim.cuh) and adds their payoff terms into 16 fp64 register
//                 accumulators -- no shared memory, no barrier inside the path loop.
//   many strikes  phase A: every thread simulates one path of a 256-path batch and parks S_T of each state in shared
//                 memory; phase B: thread t owns strike (t % n_strikes) and slice (t / n_strikes) of the batch and adds
//                 the payoff terms of that strike (per-batch partials in the state's precision, folded into fp64).
// At the end the threads of a CTA are folded through shared memory, one partial per (CTA, strike) goes to scratch and
// the last CTA to arrive adds the partials in CTA order, so a given launch geometry is bitwise reproducible.
#include "prep.cuh"

namespace b200mc {

constexpr int EU_THREADS = 256;
// Register budget: measured on B200, letting ptxas take ~100 registers (2 CTAs = 4 warps per sub-partition) beats every
// capped variant (min blocks 4/5/6: 1.58 / 1.59 / 1.57e12 vs 1.69e12 path-steps/s) and two paths per thread (1.60e12):
// the kernel is XU/MIO-queue bound and more resident warps only deepen that queue.
#ifndef EU_MIN_BLOCKS
#define EU_MIN_BLOCKS 1
#endif
#ifndef EU_PARK
#define EU_PARK 1
#endif
#ifndef EU_MIN_BLOCKS_SV
#define EU_MIN_BLOCKS_SV 1     // resident CTAs per SM asked of ptxas for the fp32 Heston / SVJ kernels without Greeks
#endif
constexpr int NACC = 16;   // doubles per strike after b200mc_sums.n

struct EuroArgs {
    ModelArgs m;
    PhiloxKey key;
    uint64_t path0;
    int64_t n_paths;
    int32_t n_steps;
    int32_t n_strikes;
    int32_t is_call;
    int32_t wld;
    int32_t fold_inside;                       // 1: the last CTA to finish adds the CTA partials; 0: k_fold_partials does
    int32_t park_off;                          // byte offset of the parked accumulators in dynamic shared memory
    double up_mul, dn_mul, rup_mul, rdn_mul;   // S_T multipliers of the spot and rate bumps
    double sigmaT;                             // sqrt(v0) T        (pathwise vega, GBM only)
    double w_scale;                            // sqrt(dt) BM_SCALE (W_T = w_scale * sum raw z)
    double strikes[EU_THREADS];                // by value: no host->device copy, no stream sync on the call path
};

template <typename R> __device__ __forceinline__ R payoff(R s, R k, bool call)
{
    return call ? rmax(s - k, (R)0) : rmax(k - s, (R)0);
}

// Adds one path's payoff terms for strike K into the accumulators (order = b200mc_sums after .n).  A = double for the
// per-thread running sums; A = float for the per-batch partials of the multi-strike fp32 kernels (at most 256 paths
// are added in fp32 before the partial is folded into the fp64 sum, so the relative rounding error stays ~1e-7).
template <int MODE, bool ANTI, bool GREEKS, typename R, typename A>
__device__ __forceinline__ void accumulate(A (&acc)[NACC], const EuroArgs &a, R K, R S0, bool call, R sa, R sb,
                                           R s_up, R s_dn, R sumz)
{
    const A da = (A)payoff<R>(sa, K, call);
    A db = (A)0, s_avg = (A)sa, pay = da;
    if constexpr (ANTI) {
        db = (A)payoff<R>(sb, K, call);
        s_avg = (A)0.5 * ((A)sa + (A)sb);
        pay = (A)0.5 * (da + db);
    }
    acc[0] += da;
    acc[2] += da * da;
    if constexpr (ANTI) {          // without the twin these three stay exactly 0 (x + 0.0 is not folded by the compiler)
        acc[1] += db;
        acc[3] += db * db;
        acc[4] += da * db;
    }
    acc[5] += s_avg;
    acc[6] += s_avg * s_avg;
    acc[7] += pay * s_avg;
    if constexpr (GREEKS) {
        const bool itm = call ? (sa > K) : (sa < K);                       // greeks.py:72,75
        if (itm) acc[8] += (A)(sa / S0);
        acc[9] += (A)payoff<R>(sa * (R)a.up_mul, K, call);
        acc[10] += (A)payoff<R>(sa * (R)a.dn_mul, K, call);
        acc[11] += (A)payoff<R>(s_up, K, call);
        acc[12] += (A)payoff<R>(s_dn, K, call);
        acc[13] += (A)payoff<R>(sa * (R)a.rup_mul, K, call);
        acc[14] += (A)payoff<R>(sa * (R)a.rdn_mul, K, call);
        if constexpr (MODE == MODE_GBM) {
            if (itm) acc[15] += (A)sa * ((A)sumz * (A)a.w_scale - (A)a.sigmaT);
        }
    }
}

// SINGLE = exactly one strike: every thread finishes its own paths, no shared-memory staging and no block barrier
// inside the path loop.  Otherwise phase A / phase B as described at the top of the file.
template <int MODE, bool ANTI, bool GREEKS, typename R, bool SINGLE, bool WIDE = false>
__global__ void __launch_bounds__(EU_THREADS, (sizeof(R) == 4 && MODE <= MODE_DETVAR) ? EU_MIN_BLOCKS
                                              : ((sizeof(R) == 4 && !GREEKS && SINGLE) ? EU_MIN_BLOCKS_SV : 1))
k_european(const __grid_constant__ EuroArgs a, const double *__restrict__ wtab_g, double *__restrict__ partials, unsigned int *counter,
           double *__restrict__ out)
{
    using L = StateLayout<ANTI, GREEKS>;
    constexpr int NS = L::NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *red = reinterpret_cast<double *>(smem_raw);                   // [256]
    double *strikes = red + EU_THREADS;                                   // [n_strikes]
    R *sT = reinterpret_cast<R *>(strikes + ((a.n_strikes + 1) & ~1));    // [NS][256]   (not SINGLE); 16-byte aligned
    R *sW = sT + (SINGLE ? 0 : NS * EU_THREADS);                          // [256]  sum of raw z (not SINGLE)
    R *wtab = sW + (SINGLE ? 0 : EU_THREADS);                             // [3][wld] (DETVAR)
    // Greeks with stochastic variance (four states per path), one strike: the 16 fp64 accumulators (32 registers) live
    // in shared memory between paths -- they are touched once per path -- and the hot loop fits 3 CTAs per SM instead
    // of 2: SVJ + Greeks 1.937 -> 1.851 ms per 2.5M x 250.  (The GBM / DETVAR kernels gain nothing from the third CTA:
    // 1.434 -> 1.447 ms, they are XU / issue bound, so they keep their accumulators in registers.)
    constexpr bool PARK = SINGLE && GREEKS && EU_PARK && MODE >= MODE_HESTON;
    double *park = reinterpret_cast<double *>(smem_raw + a.park_off);     // [NACC][256]  (PARK)

    const int tid = threadIdx.x;
    for (int i = tid; i < a.n_strikes; i += EU_THREADS) strikes[i] = a.strikes[i];
    if constexpr (MODE == MODE_DETVAR)
        for (int i = tid; i < 3 * a.wld; i += EU_THREADS) wtab[i] = (R)wtab_g[i];
    __syncthreads();

    const int ks = a.n_strikes;
    const int nslices = EU_THREADS / ks;
    const int my_k = tid % ks, my_slice = tid / ks;
    const bool worker = my_slice < nslices;
    const bool call = a.is_call != 0;
    const R K = (R)strikes[my_k];
    const R S0 = (R)a.m.S0;

    double acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.0;

    if constexpr (SINGLE) {
        if constexpr (PARK) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) park[j * EU_THREADS + tid] = 0.0;
        }
        for (int64_t i = (int64_t)blockIdx.x * EU_THREADS + tid; i < a.n_paths; i += (int64_t)gridDim.x * EU_THREADS) {
            R xT[NS], vT[NS], sumz;
            simulate_path<MODE, ANTI, GREEKS, R, NoRec, WIDE, sizeof(R) == 4>(a.m, a.key, a.path0 + (uint64_t)i, a.n_steps, wtab, a.wld, xT,
                                                              vT, sumz, NoRec());
            R sv[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) sv[k] = S0 * rexp(xT[k]);
            if constexpr (PARK) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) acc[j] = park[j * EU_THREADS + tid];
            }
            accumulate<MODE, ANTI, GREEKS, R, double>(acc, a, K, S0, call, sv[0], ANTI ? sv[NS > 1 ? 1 : 0] : (R)0,
                                              GREEKS ? sv[L::UP_IDX < NS ? L::UP_IDX : 0] : (R)0,
                                              GREEKS ? sv[L::DN_IDX < NS ? L::DN_IDX : 0] : (R)0, sumz);
            if constexpr (PARK) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) park[j * EU_THREADS + tid] = acc[j];
            }
        }
        if constexpr (PARK) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[j] = park[j * EU_THREADS + tid];
        }
    } else {
        for (int64_t base = (int64_t)blockIdx.x * EU_THREADS; base < a.n_paths;
             base += (int64_t)gridDim.x * EU_THREADS) {
            // ---- phase A: one path per thread ------------------------------------------------------------
            const int64_t i = base + tid;
            if (i < a.n_paths) {
                R xT[NS], vT[NS], sumz;
                simulate_path<MODE, ANTI, GREEKS, R>(a.m, a.key, a.path0 + (uint64_t)i, a.n_steps, wtab, a.wld, xT,
                                                     vT, sumz, NoRec());
#pragma unroll
                for (int k = 0; k < NS; ++k) sT[k * EU_THREADS + tid] = S0 * rexp(xT[k]);
                if constexpr (GREEKS && MODE == MODE_GBM) sW[tid] = sumz;
            }
            __syncthreads();
            // ---- phase B: strike-major payoff sums --------------------------------------------------------
            const int64_t left = a.n_paths - base;
            const int nvalid = left < EU_THREADS ? (int)left : EU_THREADS;
            if (worker) {
                R part[NACC];                            // per-batch partial in the state's own precision
#pragma unroll
                for (int j = 0; j < NACC; ++j) part[j] = (R)0;
                for (int p = my_slice; p < nvalid; p += nslices) {
                    accumulate<MODE, ANTI, GREEKS, R, R>(part, a, K, S0, call, sT[p], ANTI ? sT[EU_THREADS + p] : (R)0,
                                                         GREEKS ? sT[L::UP_IDX * EU_THREADS + p] : (R)0,
                                                         GREEKS ? sT[L::DN_IDX * EU_THREADS + p] : (R)0,
                                                         (GREEKS && MODE == MODE_GBM) ? sW[p] : (R)0);
                }
#pragma unroll
                for (int j = 0; j < NACC; ++j) acc[j] += (double)part[j];
            }
            __syncthreads();
        }
    }

    // ---- fold the CTA's threads -> one partial per (CTA, strike) ---------------------------------------------
    // (a tree, not one thread walking 256 slots: that walk was 35 us of serial shared-memory latency per CTA -- the
    // whole duration of a small launch and 2.4 % of a 10M-path one)
    __shared__ bool is_last;
    if constexpr (SINGLE) {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            double v = acc[j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) red[warp * NACC + j] = v;
        }
        __syncthreads();
        if (tid < NACC) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < EU_THREADS / 32; ++w) t += red[w * NACC + tid];
            partials[(size_t)blockIdx.x * NACC + tid] = t;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            red[tid] = worker ? acc[j] : 0.0;
            __syncthreads();
            for (int st = 128; st >= 1; st >>= 1) {            // halving over the slices of a strike
                if (worker && my_slice < st && my_slice + st < nslices) red[tid] += red[tid + st * ks];
                __syncthreads();
            }
            if (tid < ks) partials[((size_t)blockIdx.x * ks + tid) * NACC + j] = red[tid];
            __syncthreads();
        }
    }
    if (!a.fold_inside) return;         // many strikes: a separate multi-CTA kernel adds the partials (k_fold_partials)
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        // The last CTA adds the per-CTA partials.  A single thread per item walking all CTAs is a chain of gridDim.x
        // dependent L2 round trips (measured: ~40 us of fixed latency per launch, visible in every small call and ~3 % of
        // a 10M-path launch).  Here LANES threads share an item, each adds a strided subset of the CTAs with the loads
        // of a subset issued back to back, and the subsets are folded in a fixed order: still bitwise reproducible
        // for a given launch geometry.
        __threadfence();
        const int items = ks * NACC;
        const int LANES = items <= 16 ? 16 : (items <= 32 ? 8 : (items <= 64 ? 4 : 2));
        for (int it0 = 0; it0 < items; it0 += EU_THREADS / LANES) {      // a single pass (items <= 128)
            const int it = it0 + tid / LANES, sub = tid % LANES;
            double s = 0.0;
            if (it < items) {
                unsigned int b = sub;
                for (; b + 3 * LANES < gridDim.x; b += 4 * LANES) {
                    const double p0 = partials[(size_t)b * items + it], p1 = partials[(size_t)(b + LANES) * items + it],
                                 p2 = partials[(size_t)(b + 2 * LANES) * items + it],
                                 p3 = partials[(size_t)(b + 3 * LANES) * items + it];
                    s += p0; s += p1; s += p2; s += p3;
                }
                for (; b < gridDim.x; b += LANES) s += partials[(size_t)b * items + it];
            }
            __syncthreads();
            red[tid] = s;
            __syncthreads();
            if (it < items && sub == 0) {
                double tot = 0.0;
                for (int q = 0; q < LANES; ++q) tot += red[tid + q];
                const int k = it / NACC, j = it % NACC;
                out[(size_t)k * (NACC + 1) + 1 + j] = tot;
                if (j == 0) out[(size_t)k * (NACC + 1)] = (double)a.n_paths;
            }
        }
        if (tid == 0) *counter = 0u;   // re-arm for the next launch on this stream
    }
}

// Sum of the CTA partials for launches with many strikes (items = n_strikes * 16 > 128): one CTA reading
// gridDim.x * items doubles is limited by the bandwidth of a single SM (4.8 MB for 64 strikes: ~70 us), so this runs as
// its own small grid, 64 items per CTA, 4 lanes per item over the CTA index, folded in a fixed order.
__global__ void __launch_bounds__(256)
k_fold_partials(const double *__restrict__ partials, int n_ctas, int items, double n_paths, double *__restrict__ out)
{
    const int it = blockIdx.x * 64 + (threadIdx.x >> 2), sub = threadIdx.x & 3;
    double s = 0.0;
    if (it < items) {
        int b = sub;
        for (; b + 12 < n_ctas; b += 16) {
            const double p0 = partials[(size_t)b * items + it], p1 = partials[(size_t)(b + 4) * items + it],
                         p2 = partials[(size_t)(b + 8) * items + it], p3 = partials[(size_t)(b + 12) * items + it];
            s += p0; s += p1; s += p2; s += p3;
        }
        for (; b < n_ctas; b += 4) s += partials[(size_t)b * items + it];
    }
    const double s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2),
                 s3 = __shfl_down_sync(0xffffffffu, s, 3);
    if (it < items && sub == 0) {
        const double tot = ((s + s1) + s2) + s3;
        const int k = it / NACC, j = it % NACC;
        out[(size_t)k * (NACC + 1) + 1 + j] = tot;
        if (j == 0) out[(size_t)k * (NACC + 1)] = n_paths;
    }
}

using EuroKernel = void (*)(const EuroArgs, const double *, double *, unsigned int *, double *);

template <int MODE, typename R, bool SINGLE> static EuroKernel pick3(bool anti, bool greeks)
{
    if (anti) return greeks ? k_european<MODE, true, true, R, SINGLE> : k_european<MODE, true, false, R, SINGLE>;
    return greeks ? k_european<MODE, false, true, R, SINGLE> : k_european<MODE, false, false, R, SINGLE>;
}
template <int MODE, typename R> static EuroKernel pick2(bool anti, bool greeks, bool single)
{
    return single ? pick3<MODE, R, true>(anti, greeks) : pick3<MODE, R, false>(anti, greeks);
}
template <typename R> static EuroKernel pick1(int mode, bool anti, bool greeks, bool single)
{
    switch (mode) {
    case MODE_GBM: return pick2<MODE_GBM, R>(anti, greeks, single);
    case MODE_DETVAR: return pick2<MODE_DETVAR, R>(anti, greeks, single);
    case MODE_HESTON: return pick2<MODE_HESTON, R>(anti, greeks, single);
    default: return pick2<MODE_SVJ, R>(anti, greeks, single);
    }
}

static int launch_european(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                           int64_t n_paths, uint64_t seed, uint64_t path_offset, const double *strikes,
                           int32_t n_strikes, int is_call, uint32_t flags, const b200mc_bumps *bumps,
                           double *out_dev /* [n_strikes][17] */)
{
    Prep pr;
    B200MC_TRY(prepare(h, p, S0, T, n_steps, n_paths, seed, flags, bumps, pr));
    if (!strikes || n_strikes <= 0) return fail(h, B200MC_EINVAL, "strikes must hold at least one strike");
    if (n_strikes > EU_THREADS) return fail(h, B200MC_EINVAL, "at most 256 strikes per launch");
    if (!out_dev) return fail(h, B200MC_EINVAL, "out is NULL");

    const bool anti = flags & B200MC_ANTITHETIC, greeks = flags & B200MC_GREEKS, fp64 = flags & B200MC_FP64;
    EuroArgs a;
    memset(&a, 0, sizeof(a));
    a.m = pr.m;
    a.key = pr.key;
    a.path0 = path_offset;
    a.n_paths = n_paths;
    a.n_steps = n_steps;
    a.n_strikes = n_strikes;
    a.is_call = is_call ? 1 : 0;
    a.wld = pr.wld;
    a.fold_inside = (n_strikes * NACC <= 128) ? 1 : 0;
    if (greeks) {
        a.up_mul = 1.0 + bumps->spot_bump;
        a.dn_mul = 1.0 - bumps->spot_bump;
        a.rup_mul = exp((bumps->r_up - p->r) * T);
        a.rdn_mul = exp((bumps->r_dn - p->r) * T);
    }
    a.sigmaT = sqrt(p->v0 > 0.0 ? p->v0 : 0.0) * T;
    a.w_scale = pr.m.sqrt_dt_s;

    const bool single = n_strikes == 1;
    EuroKernel kern = fp64 ? pick1<double>(pr.mode, anti, greeks, single) : pick1<float>(pr.mode, anti, greeks, single);
    if (flags & B200MC_WIDE_RNG) {           // validation twin of the generator: constant variance, one strike, no Greeks
        if (pr.mode != MODE_GBM || !single || greeks)
            return fail(h, B200MC_EINVAL, "B200MC_WIDE_RNG is the validation twin of the GBM stream: constant variance, one strike, no B200MC_GREEKS");
        if (fp64) kern = anti ? k_european<MODE_GBM, true, false, double, true, true> : k_european<MODE_GBM, false, false, double, true, true>;
        else kern = anti ? k_european<MODE_GBM, true, false, float, true, true> : k_european<MODE_GBM, false, false, float, true, true>;
    }
    const int ns = 1 + (anti ? 1 : 0) + (greeks ? 2 : 0);
    const size_t rsz = fp64 ? 8 : 4;
    size_t smem = (size_t)(EU_THREADS + ((n_strikes + 1) & ~1)) * 8 + (single ? 0 : (size_t)(ns + 1) * EU_THREADS * rsz);
    if (pr.mode == MODE_DETVAR) smem += (size_t)3 * pr.wld * rsz;
    smem = (smem + 15) & ~(size_t)15;
    a.park_off = (int32_t)smem;
    if (EU_PARK && single && greeks && pr.mode >= MODE_HESTON) smem += (size_t)NACC * EU_THREADS * 8;
    if (smem > 200 * 1024) return fail(h, B200MC_EINVAL, "too many steps for the deterministic-variance tables");
    // dynamic-smem opt-in and occupancy are properties of (kernel, smem): query once, then reuse (small calls are
    // latency bound, calibration issues 1e4-1e5 of them)
    int occ = 0;
    for (int i = 0; i < h->n_occ; ++i)
        if (h->occ_kern[i] == (const void *)kern && h->occ_smem[i] == smem) occ = h->occ_val[i];
    if (occ == 0) {
        // the attribute is a CAP on the dynamic shared memory of later launches, not a reservation: raise it to the
        // device limit once and never lower it (setting it to this call's size would make a later, larger launch of
        // the same kernel fail with "invalid argument" when its occupancy comes out of the cache)
        B200MC_CUDA(h, cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            h->smem_optin - 1024));      // minus the kernel's few static bytes
        B200MC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)kern, EU_THREADS, smem));
        if (occ < 1) return fail(h, B200MC_ECUDA, "fused kernel does not fit on an SM");
        const int slot = h->n_occ < 64 ? h->n_occ++ : 63;
        h->occ_kern[slot] = (const void *)kern;
        h->occ_smem[slot] = smem;
        h->occ_val[slot] = occ;
    }
    // One persistent wave: CTAs stride over the paths.  (Measured: 4 or 16 waves of smaller CTAs are not faster --
    // 1.64e12 -> 1.64e12 / 1.35e12 path-steps/s -- the final per-CTA fold is what grows.)
    const int64_t need = (n_paths + EU_THREADS - 1) / EU_THREADS;
    int64_t grid = (int64_t)h->sm_count * occ;
    if (grid > need) grid = need;

    // scratch: [wtab 3*wld][partials grid*n_strikes*NACC]
    memcpy(a.strikes, strikes, (size_t)n_strikes * 8);
    const size_t off_p = pr.wtab.size() * 8;
    const size_t total = off_p + (size_t)grid * n_strikes * NACC * 8;
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, total));
    if (!pr.wtab.empty()) {     // deterministic-variance tables (rare mode): staged through the pinned bounce buffer
        B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, off_p > 4096 ? off_p : 4096, true));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));      // the buffer may still feed the previous launch's copy
        memcpy(h->h_pinned, pr.wtab.data(), off_p);
        B200MC_CUDA(h, cudaMemcpyAsync(h->d_scratch, h->h_pinned, off_p, cudaMemcpyHostToDevice, h->stream));
    }
    char *sc = (char *)h->d_scratch;
    kern<<<(unsigned)grid, EU_THREADS, smem, h->stream>>>(a, (const double *)sc, (double *)(sc + off_p), h->d_counter,
                                                          out_dev);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (!a.fold_inside) {
        const int items = n_strikes * NACC;
        k_fold_partials<<<(items + 63) / 64, 256, 0, h->stream>>>((const double *)(sc + off_p), (int)grid, items,
                                                                 (double)n_paths, out_dev);
        B200MC_CUDA(h, cudaGetLastError());
        h->launches += 1;
    }
    return 0;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_price_european_async(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                           int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                                           const double *strikes, int32_t n_strikes, int is_call, uint32_t flags,
                                           const b200mc_bumps *bumps, b200mc_sums *out_dev)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    static_assert(sizeof(b200mc_sums) == (NACC + 1) * sizeof(double), "b200mc_sums layout");
    return launch_european(h, p, S0, T, n_steps, n_paths, seed, path_offset, strikes, n_strikes, is_call, flags,
                           bumps, reinterpret_cast<double *>(out_dev));
}

extern "C" int b200mc_price_european(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                     int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                                     const double *strikes, int32_t n_strikes, int is_call, uint32_t flags,
                                     const b200mc_bumps *bumps, b200mc_sums *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out) return fail(h, B200MC_EINVAL, "out is NULL");
    if (n_strikes <= 0 || n_strikes > EU_THREADS) return fail(h, B200MC_EINVAL, "n_strikes must be in [1, 256]");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)n_strikes * sizeof(b200mc_sums);
    B200MC_TRY(ensure(h, &h->d_result, &h->result_bytes, bytes));
    B200MC_TRY(ensure(h, &h->h_result, &h->h_result_bytes, bytes, true));
    double *mapped = result_mapped_ptr(h, bytes);      // common.cuh: the kernel writes the pinned landing buffer itself
    B200MC_TRY(launch_european(h, p, S0, T, n_steps, n_paths, seed, path_offset, strikes, n_strikes, is_call, flags,
                               bumps, mapped ? mapped : reinterpret_cast<double *>(h->d_result)));
    // device -> PINNED host (a pageable destination makes the copy a staged, much slower one), then a plain memcpy
    if (!mapped) B200MC_CUDA(h, cudaMemcpyAsync(h->h_result, h->d_result, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(out, h->h_result, bytes);
    return 0;
}
