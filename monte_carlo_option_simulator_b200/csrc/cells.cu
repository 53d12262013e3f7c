// cells.cu -- many independent pricing problems ("cells") in ONE launch over a (cell x path) grid.
//
// The reference prices scenario ladders and optimiser populations as loops of MonteCarloEngine.price():
//   StressTestEngine.spot_shock_ladder / vol_shock_ladder / jump_scenario   engine/risk.py:33-111   (13 calls)
//   HedgingBacktest.run_backtest, premium of every scenario                 engine/risk.py:264-273  (num_scenarios calls)
//   _heston_objective / _svj_objective inside differential_evolution        engine/calibration.py:78-89,119-130
// Each of those calls is small (5e4 - 2e5 paths x 10 - 250 steps: 10 - 40 us of GPU work), so a loop of launches is
// bound by launch latency and by the tail of every launch.  Here a cell carries its own model parameters, spot,
// maturity, step count, path count, seed and strikes; a CTA looks up its cell, copies the cell's constants into shared
// memory and runs the same per-path simulator (sim.cuh) and the same draw layout as k_european, so a cell's sums equal
// those of a b200mc_price_european call with the cell's arguments up to the order of the fp64 additions.
//
// Grid: cells of one mode (GBM / DETVAR / HESTON / SVJ) form a group; every cell of a group gets `cpc` CTAs, CTA `part`
// of a cell takes the 256-path batches part, part + cpc, ...  Partials [CTA][strike][8] go to scratch and a second small
// kernel (one warp per (cell, strike, sum)) adds them in CTA order: bitwise reproducible for a given geometry.
#include <algorithm>
#include <vector>

#include "prep.cuh"

namespace b200mc {

constexpr int CL_THREADS = 256;
constexpr int CL_NACC = 8;      // sum_a sum_b sum_aa sum_bb sum_ab sum_s sum_ss sum_ps (b200mc_sums, price part)
constexpr int CL_NOUT = 17;     // doubles per b200mc_sums

struct CellDev {
    ModelArgs m;
    PhiloxKey key;
    uint64_t path0;
    int64_t n_paths;
    int32_t n_steps, is_call, wld, tab_off;     // tab_off: offset (doubles) of the cell's DETVAR weight row
};
static_assert(sizeof(CellDev) % 8 == 0, "CellDev is copied as 64-bit words");

template <typename R> __device__ __forceinline__ R cl_payoff(R s, R k, bool call)
{
    return call ? rmax(s - k, (R)0) : rmax(k - s, (R)0);
}

// Same terms and the same order as accumulate() of european.cu without the Greek sums.
template <bool ANTI, typename R, typename A>
__device__ __forceinline__ void cl_accumulate(A (&acc)[CL_NACC], R K, bool call, R sa, R sb)
{
    const A da = (A)cl_payoff<R>(sa, K, call);
    A db = (A)0, s_avg = (A)sa, pay = da;
    if constexpr (ANTI) {
        db = (A)cl_payoff<R>(sb, K, call);
        s_avg = (A)0.5 * ((A)sa + (A)sb);
        pay = (A)0.5 * (da + db);
    }
    acc[0] += da;
    acc[2] += da * da;
    if constexpr (ANTI) {
        acc[1] += db;
        acc[3] += db * db;
        acc[4] += da * db;
    }
    acc[5] += s_avg;
    acc[6] += s_avg * s_avg;
    acc[7] += pay * s_avg;
}

template <int MODE, bool ANTI, typename R, bool SINGLE>
__global__ void __launch_bounds__(CL_THREADS)
k_cells(const CellDev *__restrict__ cells, const int32_t *__restrict__ order, const double *__restrict__ strikes_g,
        const double *__restrict__ wtab_g, int ks, int cpc, int wld_max, double *__restrict__ partials)
{
    using L = StateLayout<ANTI, false>;
    constexpr int NS = L::NS;
    __shared__ __align__(16) CellDev c;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *red = reinterpret_cast<double *>(smem_raw);                   // [8 warps][CL_NACC] or [256]
    double *strikes = red + CL_THREADS;                                   // [ks]
    R *sT = reinterpret_cast<R *>(strikes + ((ks + 1) & ~1));             // [NS][256]   (not SINGLE)
    R *wtab = sT + (SINGLE ? 0 : NS * CL_THREADS);                        // [wld]       (DETVAR)

    const int tid = threadIdx.x;
    const int slot = blockIdx.x / cpc, part = blockIdx.x % cpc;
    const int cell = order[slot];
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(cells + cell);
        uint64_t *dst = reinterpret_cast<uint64_t *>(&c);
        for (int i = tid; i < (int)(sizeof(CellDev) / 8); i += CL_THREADS) dst[i] = src[i];
    }
    for (int i = tid; i < ks; i += CL_THREADS) strikes[i] = strikes_g[(size_t)cell * ks + i];
    __syncthreads();
    if constexpr (MODE == MODE_DETVAR) {
        for (int i = tid; i < wld_max; i += CL_THREADS) wtab[i] = i < c.wld ? (R)wtab_g[c.tab_off + i] : (R)0;
        __syncthreads();
    }

    const int nslices = CL_THREADS / ks;
    const int my_k = tid % ks, my_slice = tid / ks;
    const bool worker = my_slice < nslices;
    const bool call = c.is_call != 0;
    const R K = (R)strikes[my_k];
    const R S0 = (R)c.m.S0;
    const int64_t n_paths = c.n_paths;

    double acc[CL_NACC];
#pragma unroll
    for (int j = 0; j < CL_NACC; ++j) acc[j] = 0.0;

    for (int64_t base = (int64_t)part * CL_THREADS; base < n_paths; base += (int64_t)cpc * CL_THREADS) {
        const int64_t i = base + tid;
        if constexpr (SINGLE) {
            if (i < n_paths) {
                R xT[NS], vT[NS], sumz;
                simulate_path<MODE, ANTI, false, R>(c.m, c.key, c.path0 + (uint64_t)i, c.n_steps, wtab, c.wld, xT, vT,
                                                    sumz, NoRec());
                const R sa = S0 * rexp(xT[0]);
                const R sb = ANTI ? S0 * rexp(xT[NS - 1]) : (R)0;
                cl_accumulate<ANTI, R, double>(acc, K, call, sa, sb);
            }
        } else {
            if (i < n_paths) {                                           // phase A: one path per thread
                R xT[NS], vT[NS], sumz;
                simulate_path<MODE, ANTI, false, R>(c.m, c.key, c.path0 + (uint64_t)i, c.n_steps, wtab, c.wld, xT, vT,
                                                    sumz, NoRec());
#pragma unroll
                for (int k = 0; k < NS; ++k) sT[k * CL_THREADS + tid] = S0 * rexp(xT[k]);
            }
            __syncthreads();
            const int64_t left = n_paths - base;
            const int nvalid = left < CL_THREADS ? (int)left : CL_THREADS;
            if (worker) {                                                // phase B: strike-major payoff sums
                R pt[CL_NACC];
#pragma unroll
                for (int j = 0; j < CL_NACC; ++j) pt[j] = (R)0;
                for (int p = my_slice; p < nvalid; p += nslices)
                    cl_accumulate<ANTI, R, R>(pt, K, call, sT[p], ANTI ? sT[CL_THREADS + p] : (R)0);
#pragma unroll
                for (int j = 0; j < CL_NACC; ++j) acc[j] += (double)pt[j];
            }
            __syncthreads();
        }
    }

    if constexpr (SINGLE) {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int j = 0; j < CL_NACC; ++j) {
            double v = acc[j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) red[warp * CL_NACC + j] = v;
        }
        __syncthreads();
        if (tid < CL_NACC) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < CL_THREADS / 32; ++w) t += red[w * CL_NACC + tid];
            partials[(size_t)blockIdx.x * CL_NACC + tid] = t;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CL_NACC; ++j) {
            red[tid] = worker ? acc[j] : 0.0;
            __syncthreads();
            for (int st = 128; st >= 1; st >>= 1) {
                if (worker && my_slice < st && my_slice + st < nslices) red[tid] += red[tid + st * ks];
                __syncthreads();
            }
            if (tid < ks) partials[((size_t)blockIdx.x * ks + tid) * CL_NACC + j] = red[tid];
            __syncthreads();
        }
    }
}

// One warp per (slot, strike, sum): lanes add the cell's CTA partials strided by 32, then a fixed-order shuffle fold.
__global__ void __launch_bounds__(256)
k_cells_fold(const double *__restrict__ partials, const CellDev *__restrict__ cells, const int32_t *__restrict__ order,
             int n_slots, int ks, int cpc, double *__restrict__ out)
{
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= n_slots * ks * CL_NACC) return;
    const int j = item % CL_NACC, k = (item / CL_NACC) % ks, slot = item / (CL_NACC * ks);
    double s = 0.0;
    for (int p = lane; p < cpc; p += 32) s += partials[(((size_t)slot * cpc + p) * ks + k) * CL_NACC + j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) {
        const int cell = order[slot];
        double *o = out + ((size_t)cell * ks + k) * CL_NOUT;
        o[1 + j] = s;
        if (j == 0) {
            o[0] = (double)cells[cell].n_paths;
#pragma unroll
            for (int q = 1 + CL_NACC; q < CL_NOUT; ++q) o[q] = 0.0;
        }
    }
}

using CellKernel = void (*)(const CellDev *, const int32_t *, const double *, const double *, int, int, int, double *);

template <int MODE, typename R> static CellKernel cl_pick2(bool anti, bool single)
{
    if (anti) return single ? k_cells<MODE, true, R, true> : k_cells<MODE, true, R, false>;
    return single ? k_cells<MODE, false, R, true> : k_cells<MODE, false, R, false>;
}
template <typename R> static CellKernel cl_pick1(int mode, bool anti, bool single)
{
    switch (mode) {
    case MODE_GBM: return cl_pick2<MODE_GBM, R>(anti, single);
    case MODE_DETVAR: return cl_pick2<MODE_DETVAR, R>(anti, single);
    case MODE_HESTON: return cl_pick2<MODE_HESTON, R>(anti, single);
    default: return cl_pick2<MODE_SVJ, R>(anti, single);
    }
}

static size_t up8(size_t x) { return (x + 7) & ~(size_t)7; }

static int launch_cells(b200mc_handle *h, const b200mc_cell *cells, int32_t n_cells, const double *strikes,
                        int32_t n_strikes, uint32_t flags, double *out_dev)
{
    if (!cells || n_cells <= 0) return fail(h, B200MC_EINVAL, "cells must hold at least one cell");
    if (!strikes || n_strikes <= 0 || n_strikes > CL_THREADS)
        return fail(h, B200MC_EINVAL, "n_strikes must be in [1, 256]");
    if (flags & B200MC_GREEKS) return fail(h, B200MC_EINVAL, "b200mc_price_cells has no Greek sums (use b200mc_price_european)");
    if (flags & B200MC_WIDE_RNG) return fail(h, B200MC_EINVAL, "B200MC_WIDE_RNG is valid for b200mc_price_european only");
    if (n_cells > (1 << 20)) return fail(h, B200MC_EINVAL, "at most 2^20 cells per call");
    const bool anti = flags & B200MC_ANTITHETIC, fp64 = flags & B200MC_FP64;

    // ---- host: constants of every cell, grouping by mode --------------------------------------------------------
    std::vector<CellDev> dev((size_t)n_cells);
    std::vector<int> mode((size_t)n_cells);
    std::vector<double> tabs;
    int64_t max_paths[4] = {0, 0, 0, 0};
    int32_t max_steps[4] = {0, 0, 0, 0};
    int wld_max = 0, count[4] = {0, 0, 0, 0};
    Prep pr;
    for (int32_t i = 0; i < n_cells; ++i) {
        const b200mc_cell &cl = cells[i];
        B200MC_TRY(prepare(h, &cl.params, cl.S0, cl.T, cl.n_steps, cl.n_paths, cl.seed, flags, nullptr, pr));
        CellDev &d = dev[(size_t)i];
        memset(&d, 0, sizeof(d));
        d.m = pr.m;
        d.key = pr.key;
        d.path0 = cl.path_offset;
        d.n_paths = cl.n_paths;
        d.n_steps = cl.n_steps;
        d.is_call = cl.is_call ? 1 : 0;
        d.wld = pr.wld;
        if (pr.mode == MODE_DETVAR) {
            d.tab_off = (int32_t)tabs.size();
            tabs.insert(tabs.end(), pr.wtab.begin(), pr.wtab.begin() + pr.wld);      // row 0: the primary state
            wld_max = std::max(wld_max, pr.wld);
            if (tabs.size() > ((size_t)1 << 28)) return fail(h, B200MC_EINVAL, "deterministic-variance tables too large");
        }
        mode[(size_t)i] = pr.mode;
        count[pr.mode] += 1;
        max_paths[pr.mode] = std::max<int64_t>(max_paths[pr.mode], cl.n_paths);
        max_steps[pr.mode] = std::max(max_steps[pr.mode], cl.n_steps);
    }
    std::vector<int32_t> order;
    order.reserve((size_t)n_cells);
    int first[5] = {0, 0, 0, 0, 0};
    for (int md = 0; md < 4; ++md) {
        first[md] = (int)order.size();
        for (int32_t i = 0; i < n_cells; ++i)
            if (mode[(size_t)i] == md) order.push_back(i);
    }
    first[4] = (int)order.size();

    // ---- one blob: [cells][order][strikes][tables] -> pinned -> device ------------------------------------------
    const size_t off_cells = 0, off_order = up8(off_cells + dev.size() * sizeof(CellDev)),
                 off_strk = up8(off_order + order.size() * 4), off_tabs = up8(off_strk + (size_t)n_cells * n_strikes * 8),
                 blob = up8(off_tabs + tabs.size() * 8) + 16;
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, blob, true));
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, blob));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));          // the bounce buffer may still feed an earlier copy
    char *hp = (char *)h->h_pinned;
    memcpy(hp + off_cells, dev.data(), dev.size() * sizeof(CellDev));
    memcpy(hp + off_order, order.data(), order.size() * 4);
    memcpy(hp + off_strk, strikes, (size_t)n_cells * n_strikes * 8);
    if (!tabs.empty()) memcpy(hp + off_tabs, tabs.data(), tabs.size() * 8);
    B200MC_CUDA(h, cudaMemcpyAsync(h->d_stage, hp, blob, cudaMemcpyHostToDevice, h->stream));
    const char *ds = (const char *)h->d_stage;

    // ---- geometry of every group, scratch for the partials ------------------------------------------------------
    const bool single = n_strikes == 1;
    const int ns = 1 + (anti ? 1 : 0);
    const size_t rsz = fp64 ? 8 : 4;
    struct Group { CellKernel kern; size_t smem; int cpc, slots; size_t poff; } g[4];
    size_t ptotal = 0;
    for (int md = 0; md < 4; ++md) {
        g[md].slots = first[md + 1] - first[md];
        if (!g[md].slots) continue;
        g[md].kern = fp64 ? cl_pick1<double>(md, anti, single) : cl_pick1<float>(md, anti, single);
        size_t smem = (size_t)(CL_THREADS + ((n_strikes + 1) & ~1)) * 8 + (single ? 0 : (size_t)ns * CL_THREADS * rsz);
        if (md == MODE_DETVAR) smem += (size_t)wld_max * rsz;
        if (smem > 200 * 1024) return fail(h, B200MC_EINVAL, "too many steps for the deterministic-variance tables");
        g[md].smem = smem;
        int occ = 0;
        for (int i = 0; i < h->n_occ; ++i)
            if (h->occ_kern[i] == (const void *)g[md].kern && h->occ_smem[i] == smem) occ = h->occ_val[i];
        if (occ == 0) {
            B200MC_CUDA(h, cudaFuncSetAttribute((const void *)g[md].kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                h->smem_optin - 2048));   // minus the static CellDev copy
            B200MC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)g[md].kern, CL_THREADS, smem));
            if (occ < 1) return fail(h, B200MC_ECUDA, "cell kernel does not fit on an SM");
            const int slot = h->n_occ < 64 ? h->n_occ++ : 63;
            h->occ_kern[slot] = (const void *)g[md].kern;
            h->occ_smem[slot] = smem;
            h->occ_val[slot] = occ;
        }
        // CTAs per cell: a CTA should carry ~2^18 path-steps (~90 us of GBM work: the launch then ends within a few
        // per cent of its ideal time whatever the mix of cell sizes) and the group should fill ~4 waves of resident
        // CTAs; never more CTAs than the cell has 256-path batches.  (A CTA costs ~2 us of prologue and fold.)
        const int64_t need = (max_paths[md] + CL_THREADS - 1) / CL_THREADS;
        const int64_t by_work = (max_paths[md] * (int64_t)max_steps[md] + (1 << 18) - 1) >> 18;
        const int64_t by_fill = ((int64_t)4 * h->sm_count * occ + g[md].slots - 1) / g[md].slots;
        int64_t cpc = std::min(need, std::max<int64_t>(std::max(by_work, by_fill), 1));
        while ((int64_t)g[md].slots * cpc > 0x7fffffff / 2) cpc = (cpc + 1) / 2;
        g[md].cpc = (int)cpc;
        g[md].poff = ptotal;
        ptotal += (size_t)g[md].slots * cpc * n_strikes * CL_NACC * 8;
    }
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, ptotal));
    for (int md = 0; md < 4; ++md) {
        if (!g[md].slots) continue;
        double *part = (double *)((char *)h->d_scratch + g[md].poff);
        const CellDev *dc = (const CellDev *)(ds + off_cells);
        const int32_t *dord = (const int32_t *)(ds + off_order) + first[md];
        g[md].kern<<<(unsigned)(g[md].slots * g[md].cpc), CL_THREADS, g[md].smem, h->stream>>>(
            dc, dord, (const double *)(ds + off_strk), (const double *)(ds + off_tabs), n_strikes, g[md].cpc, wld_max, part);
        B200MC_CUDA(h, cudaGetLastError());
        const int64_t items = (int64_t)g[md].slots * n_strikes * CL_NACC;
        k_cells_fold<<<(unsigned)((items + 7) / 8), 256, 0, h->stream>>>(part, dc, dord, g[md].slots, n_strikes, g[md].cpc,
                                                                         out_dev);
        B200MC_CUDA(h, cudaGetLastError());
        h->launches += 2;
    }
    return 0;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_price_cells(b200mc_handle *h, const b200mc_cell *cells, int32_t n_cells, const double *strikes,
                                  int32_t n_strikes, uint32_t flags, int on_device, b200mc_sums *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out) return fail(h, B200MC_EINVAL, "out is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    static_assert(sizeof(b200mc_sums) == CL_NOUT * sizeof(double), "b200mc_sums layout");
    if (on_device) return launch_cells(h, cells, n_cells, strikes, n_strikes, flags, reinterpret_cast<double *>(out));
    if (n_cells <= 0 || n_strikes <= 0 || n_strikes > CL_THREADS || n_cells > (1 << 20))
        return fail(h, B200MC_EINVAL, "n_cells must be in [1, 2^20] and n_strikes in [1, 256]");
    const size_t bytes = (size_t)n_cells * n_strikes * sizeof(b200mc_sums);
    B200MC_TRY(ensure(h, &h->d_result, &h->result_bytes, bytes));
    B200MC_TRY(ensure(h, &h->h_result, &h->h_result_bytes, bytes, true));
    double *mapped = result_mapped_ptr(h, bytes);      // common.cuh
    B200MC_TRY(launch_cells(h, cells, n_cells, strikes, n_strikes, flags,
                            mapped ? mapped : reinterpret_cast<double *>(h->d_result)));
    if (!mapped) B200MC_CUDA(h, cudaMemcpyAsync(h->h_result, h->d_result, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(out, h->h_result, bytes);
    return 0;
}
