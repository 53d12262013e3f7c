// paths.cu -- the path-storing modes of the fused simulator.
//
//   b200mc_simulate_terminal  S_T (+ antithetic twin, + v_T) per path: the vectors MonteCarloEngine.price reduces
//                             (engine/monte_carlo.py:310-324) and the terminal P&L input of compute_risk_metrics
//                             (engine/risk.py:117).  One thread per path, consecutive threads write consecutive
//                             elements.
//   b200mc_generate_paths     the full matrix [n_paths, n_steps + 1] of MonteCarloEngine.get_sample_paths
//                             (engine/monte_carlo.py:452-471 via :215-217,241) at scale.  HBM-write bound:
//                             4 or 8 bytes per path-step, written exactly once.  A warp owns 32 consecutive paths
//                             and a 32x32 tile in shared memory; each lane fills its own row as the steps come out
//                             of the recurrence, then the warp writes the tile out row by row so that every store
//                             instruction covers 128 (fp32) or 256 (fp64) contiguous bytes.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "prep.cuh"

namespace b200mc {

constexpr int PT_THREADS = 256;
#ifndef B200MC_PATHS_DEFAULT_POLY
#define B200MC_PATHS_DEFAULT_POLY 1
#endif
#ifndef B200MC_PATHS_DEFAULT_MINB
#define B200MC_PATHS_DEFAULT_MINB 4
#endif
constexpr float LOG2E_F = 1.4426950408889634f;

struct PathArgs {
    ModelArgs m;
    PhiloxKey key;
    uint64_t path0;
    int64_t n_paths;
    int64_t ld;
    int32_t n_steps;
    int32_t wld;
};

// ---------------------------------------------------------------------------------------------- terminal
template <int MODE, bool ANTI, typename R, typename O>
__global__ void __launch_bounds__(PT_THREADS)
k_terminal(const __grid_constant__ PathArgs a, const double *__restrict__ wtab_g, O *__restrict__ S_T,
           O *__restrict__ S_T_anti, O *__restrict__ v_T)
{
    using L = StateLayout<ANTI, false>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *wtab = reinterpret_cast<R *>(smem_raw);
    if constexpr (MODE == MODE_DETVAR) {
        for (int i = threadIdx.x; i < 3 * a.wld; i += PT_THREADS) wtab[i] = (R)wtab_g[i];
        __syncthreads();
    }
    const R S0 = (R)a.m.S0;
    for (int64_t i = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x; i < a.n_paths;
         i += (int64_t)gridDim.x * PT_THREADS) {
        R xT[L::NS], vT[L::NS], sumz;
        simulate_path<MODE, ANTI, false, R>(a.m, a.key, a.path0 + (uint64_t)i, a.n_steps, wtab, a.wld, xT, vT, sumz,
                                            NoRec());
        if (S_T) S_T[i] = (O)(S0 * rexp(xT[0]));
        if constexpr (ANTI) {
            if (S_T_anti) S_T_anti[i] = (O)(S0 * rexp(xT[1]));
        }
        if (v_T) {
            if constexpr (MODE == MODE_GBM) v_T[i] = (O)a.m.v0[0];
            else if constexpr (MODE == MODE_DETVAR) v_T[i] = (O)a.m.v0[1];   // host stores the final variance here
            else v_T[i] = (O)vT[0];
        }
    }
}

// ---------------------------------------------------------------------------------------------- path store, Heston / SVJ
// Stochastic variance makes a path sequential in time, so one lane owns one path for all its steps (the time-parallel
// TMA tile of the deterministic-variance kernel is not available).  Rows of the reference layout start on arbitrary
// element boundaries; storing fixed column tiles row by row leaves every 128-byte store straddling two lines (measured
// 0.95 TB/s).  Here every row keeps a ring of two 64-byte windows in shared memory, positioned by the row's OWN
// misalignment m_r = (address of the row start / sizeof(O)) mod A (A = elements per 64 bytes): window w of row r holds
// the columns whose global addresses fall into the w-th 64-byte piece of that row.  After every A produced columns all
// 32 rows have completed the same window index, and the warp writes them out: every store instruction covers full,
// ALIGNED 32-byte sectors only (2 rows per instruction for fp32 output, 4 for fp64).
template <int MODE, typename R, typename O> struct AlignedRec {
    static constexpr bool enabled = true;
    static constexpr int A = 64 / (int)sizeof(O);       // window = 64 bytes = two full sectors (keeps the rings small)
    O *ring;             // this warp's [32][pitch] rings (2 A elements each + padding)
    int pitch, ncol, lane, m;    // m = misalignment of MY row, in elements
    R S0;

    __device__ __forceinline__ O &slot(int c) const { return ring[lane * pitch + ((m + c) & (2 * A - 1))]; }

    // write window `w` (columns [w A - m_r, (w + 1) A - m_r) of row r) of all 32 rows.  `base` is the 64-byte aligned
    // address this lane's OWN row is measured from (row start minus m elements); rows beyond n_rows carry base = nullptr.
    O *base;
    __device__ __noinline__ void flush(int w) const
    {
        __syncwarp();
        constexpr int RPI = 32 / A;                     // rows per store instruction (2 for fp32, 4 for fp64)
        const int j = lane % A, sub = lane / A;
        const unsigned long long mine = (unsigned long long)(uintptr_t)base;
        const int off = w * A + j;
#pragma unroll 4
        for (int it = 0; it < 32 / RPI; ++it) {
            const int r = it * RPI + sub;
            const int mr = __shfl_sync(0xffffffffu, m, r);
            O *rb = reinterpret_cast<O *>((uintptr_t)__shfl_sync(0xffffffffu, mine, r));
            const int c = off - mr;
            if (rb != nullptr && c >= 0 && c < ncol) rb[off] = ring[r * pitch + ((w & 1) * A + j)];
        }
        __syncwarp();
    }
    __device__ __forceinline__ void operator()(int s, R x) const
    {
        R S;
        if constexpr (sizeof(R) == 4) S = S0 * exp2f(x * LOG2E_F);
        else S = S0 * exp(x);
        const int c = s + 1;
        slot(c) = (O)S;
        if (((c + 1) & (A - 1)) == 0) flush((c + 1) / A - 1);       // (uniform across the warp: every lane is at column c)
        if (c == ncol - 1) {        // last column: rows still hold a partial window k = floor(ncol / A), and those with a
            const int w = ncol / A; // large misalignment also the start of window k + 1 (never k - 1: it shares a ring
            flush(w);               // half with k + 1 and has been written out already)
            flush(w + 1);
        }
    }
};

template <int MODE, typename R, typename O>
__global__ void __launch_bounds__(PT_THREADS)
k_paths(const __grid_constant__ PathArgs a, const double *__restrict__ wtab_g, const double *__restrict__ dtab_g,
        O *__restrict__ out)
{
    using L = StateLayout<false, false>;
    using Rec = AlignedRec<MODE, R, O>;
    static_assert(MODE == MODE_HESTON || MODE == MODE_SVJ, "deterministic variance has its own kernels");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int pitch = 2 * Rec::A + (((a.ld & 1) != 0) ? 0 : 1);                 // (pitch + ld) odd: conflict-free fills
    O *rings = reinterpret_cast<O *>(smem_raw);                                 // [8][32 * pitch]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_groups = (a.n_paths + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * (PT_THREADS / 32) + warp; g < n_groups;
         g += (int64_t)gridDim.x * (PT_THREADS / 32)) {
        Rec rec;
        rec.ring = rings + (size_t)warp * 32 * pitch;
        rec.pitch = pitch;
        rec.ncol = a.n_steps + 1;
        rec.lane = lane;
        rec.S0 = (R)a.m.S0;
        int64_t me = g * 32 + lane;
        rec.m = (int)((((uintptr_t)out / sizeof(O)) + (uint64_t)me * (uint64_t)a.ld) % (uint64_t)Rec::A);
        rec.base = me < a.n_paths ? out + (size_t)me * a.ld - rec.m : nullptr;
        if (me >= a.n_paths) me = a.n_paths - 1;   // idle lanes shadow the last path so the warp stays convergent
        __syncwarp();
        rec.slot(0) = (O)a.m.S0;                                                // column 0 = S0   (:217)
        R xT[L::NS], vT[L::NS], sumz;
        simulate_path<MODE, false, false, R>(a.m, a.key, a.path0 + (uint64_t)me, a.n_steps, nullptr, a.wld, xT, vT, sumz,
                                             rec);
    }
}

// ---------------------------------------------------------------------------------------------- path store, GBM / DETVAR
// Dedicated kernel for deterministic variance (the BASELINE path-store configuration): no per-step branching.
// Column c of the matrix is y_c (y_0 = S0, y_{s+1} = S after step s).  A warp owns 32 consecutive paths and walks
// the columns in tiles of 32: slot 0 of a tile is the value carried over from the previous tile, slots 1..31 and
// the next carry come out of four Philox calls (32 normals per lane).  The tile is then written out row by row.
//   scalar flush  one row per store instruction: 32 lanes x sizeof(O) contiguous bytes
//   VEC flush     (fp32 output, ld % 4 == 0, 16-byte aligned base) the tile is kept as float4 groups with an
//                 XOR swizzle (group g of row r lives at g ^ (r & 7)), filled with 128-bit shared stores and
//                 written with 128-bit global stores, 4 rows (4 x 128 contiguous bytes) per instruction.
template <bool TAB, typename R, typename O, bool VEC>
__global__ void __launch_bounds__(PT_THREADS)
k_paths_det(const __grid_constant__ PathArgs a, const double *__restrict__ wtab_g, const double *__restrict__ dtab_g,
            O *__restrict__ out)
{
    static_assert(!VEC || sizeof(O) == 4, "the 128-bit flush is for fp32 output");
    constexpr int NW = PT_THREADS / 32;
    constexpr R UNIT = sizeof(R) == 4 ? (R)1.4426950408889634 : (R)1;     // fp32 carries log2(S/S0), fp64 ln(S/S0)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    O *tiles = reinterpret_cast<O *>(smem_raw);                                     // [NW][32*33] (VEC: [NW][32*32])
    R *w2 = reinterpret_cast<R *>(tiles + NW * 32 * 33);                            // [wld + 32]  TAB only
    R *d2 = w2 + a.wld + 32;                                                        // [wld + 32]
    if constexpr (TAB) {
        for (int i = threadIdx.x; i < a.wld + 32; i += PT_THREADS) {
            const bool in = i < a.n_steps;
            w2[i] = in ? (R)(wtab_g[i] * (double)UNIT) : (R)0;
            d2[i] = in ? (R)((dtab_g[i] - (i ? dtab_g[i - 1] : 0.0)) * (double)UNIT) : (R)0;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    O *tile = tiles + warp * 32 * 33;
    const R S0 = (R)a.m.S0;
    const R wc = (R)(a.m.x_w[0] * (double)UNIT), dc = (R)(a.m.step_drift[0] * (double)UNIT);
    const int ncol = a.n_steps + 1;
    const int ntiles = (ncol + 31) >> 5;
    const int64_t n_groups = (a.n_paths + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * NW + warp; g < n_groups; g += (int64_t)gridDim.x * NW) {
        const int64_t path0 = g * 32;
        int64_t me = path0 + lane;
        if (me >= a.n_paths) me = a.n_paths - 1;          // idle lanes shadow the last path (never stored)
        const uint64_t path = a.path0 + (uint64_t)me;
        const uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32);
        const int rows = (int)((a.n_paths - path0) < 32 ? (a.n_paths - path0) : 32);
        R x = (R)0;
        O carry = (O)a.m.S0;
        for (int k = 0; k < ntiles; ++k) {
            O vals[32];
            vals[0] = carry;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const U4 u = philox4x32_10(c0, c1, (uint32_t)(4 * k + b), B200MC_STREAM_GBM, a.key);
                const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const BM2 bm = box_muller_word(ww[t]);
                    const int s = 32 * k + 8 * b + 2 * t;          // step index of the first normal of the pair
                    R wa, wb, da, db;
                    if constexpr (TAB) { wa = w2[s]; wb = w2[s + 1]; da = d2[s]; db = d2[s + 1]; }
                    else { wa = wb = wc; da = db = dc; }
                    x = (x + da) + wa * (R)bm.rc;
                    O ya, yb;
                    if constexpr (sizeof(R) == 4) ya = (O)(S0 * ex2_approx((float)x)); else ya = (O)(S0 * exp(x));
                    x = (x + db) + wb * (R)bm.rs;
                    if constexpr (sizeof(R) == 4) yb = (O)(S0 * ex2_approx((float)x)); else yb = (O)(S0 * exp(x));
                    const int idx = 8 * b + 2 * t + 1;             // tile slots idx, idx + 1
                    vals[idx] = ya;
                    if (idx + 1 < 32) vals[idx + 1] = yb; else carry = yb;
                }
            }
            const int col0 = 32 * k;
            const int ncv = ncol - col0 < 32 ? ncol - col0 : 32;   // valid columns of this tile
            __syncwarp();
            if constexpr (VEC) {
                float4 *t4 = reinterpret_cast<float4 *>(tile);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    t4[lane * 8 + (q ^ (lane & 7))] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
                __syncwarp();
                const int gq = lane & 7;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = 4 * it + (lane >> 3);
                    const float4 v = t4[r * 8 + (gq ^ (r & 7))];
                    float *dst = reinterpret_cast<float *>(out) + (size_t)(path0 + r) * a.ld + col0 + 4 * gq;
                    if (r < rows) {
                        if (4 * gq + 3 < ncv) *reinterpret_cast<float4 *>(dst) = v;
                        else {
                            if (4 * gq + 0 < ncv) dst[0] = v.x;
                            if (4 * gq + 1 < ncv) dst[1] = v.y;
                            if (4 * gq + 2 < ncv) dst[2] = v.z;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) tile[lane * 33 + q] = vals[q];
                __syncwarp();
                O *dst = out + (size_t)path0 * a.ld + col0 + lane;
                if (rows == 32 && ncv == 32) {
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) { *dst = tile[r * 33 + lane]; dst += a.ld; }
                } else {
                    for (int r = 0; r < rows; ++r) { if (lane < ncv) *dst = tile[r * 33 + lane]; dst += a.ld; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- path store, TMA flush
// The reference layout [n_paths, n_steps + 1] (ld == n_steps + 1, e.g. 251 columns) has rows that start on arbitrary
// 4-byte boundaries, so row-wise 128-byte stores straddle cache lines and leave partial sectors (measured: 1.96 TB/s
// against 3.3 TB/s for 1024-byte rows).  But the rows of 32 CONSECUTIVE paths are one contiguous byte range that
// starts on a 128-byte boundary (32 * ncol * 4 bytes = ncol full lines).  This kernel therefore gives a CTA 32 paths
// and splits the TIME axis across its warps -- the counter-based generator can start anywhere: warp c, lane p draws
// steps [32c, 32c + 32) of path p.  Chunk totals meet in shared memory (the only cross-warp dependence of a path),
// every thread turns its chunk-local log returns into spots and drops them into the CTA's tile, which has exactly
// the global layout (flat [32][ncol]; the fill is bank-conflict free because ncol is odd for even step counts and
// lanes walk different rows), and ONE thread hands the whole tile to the TMA engine: a single cp.async.bulk
// shared -> global copy of 32 * ncol * sizeof(O) bytes (UBLKCP), fully aligned, overlapped with the next tile's
// Philox work.  No per-row store instructions, no partial sectors.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// DEG > 0 (fp32 state, constant variance): the spot is carried MULTIPLICATIVELY, S_{t+1} = S_t + S_t expm1(delta_t) with
// expm1 a degree-DEG Taylor polynomial on the FMA pipe (|delta| is small: sigma sqrt(dt) * 5.65 + |drift dt|, the host
// picks DEG from that bound and falls back to the MUFU form when it is too large) -- no ex2 per stored value, i.e. 2
// instead of 3 MUFU per path-step.
template <bool TAB, typename R, typename O, int MAXT, int DEG, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB)
k_paths_tma(const __grid_constant__ PathArgs a, const double *__restrict__ wtab_g, const double *__restrict__ dtab_g,
            O *__restrict__ out)
{
    constexpr R UNIT = sizeof(R) == 4 ? (R)1.4426950408889634 : (R)1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ncol = a.n_steps + 1;
    const int NCH = blockDim.x >> 5;                                                // chunks of 32 steps = warps
    O *tile = reinterpret_cast<O *>(smem_raw);                                      // [32][ncol], the global layout
    const size_t tile_bytes = ((size_t)32 * ncol * sizeof(O) + 127) & ~(size_t)127;
    R *totals = reinterpret_cast<R *>(smem_raw + tile_bytes);                       // [NCH][32]
    R *w2 = totals + NCH * 32;                                                      // [wld + 32]  TAB only
    R *d2 = w2 + a.wld + 32;
    if constexpr (TAB) {
        for (int i = threadIdx.x; i < a.wld + 32; i += blockDim.x) {
            const bool in = i < a.n_steps;
            w2[i] = in ? (R)(wtab_g[i] * (double)UNIT) : (R)0;
            d2[i] = in ? (R)((dtab_g[i] - (i ? dtab_g[i - 1] : 0.0)) * (double)UNIT) : (R)0;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    const R S0 = (R)a.m.S0;
    static_assert(DEG == 0 || (!TAB && sizeof(R) == 4), "the polynomial form is for the constant-variance fp32 state");
    const R wc = (R)(a.m.x_w[0] * (double)(DEG ? (R)1 : UNIT)), dc = (R)(a.m.step_drift[0] * (double)(DEG ? (R)1 : UNIT));
    const int64_t n_groups = (a.n_paths + 31) / 32;
    const bool dst_aligned = ((uintptr_t)out & 15) == 0;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int64_t path0 = g * 32;
        int64_t me = path0 + lane;
        if (me >= a.n_paths) me = a.n_paths - 1;                                    // idle lanes shadow the last path
        const uint64_t path = a.path0 + (uint64_t)me;
        const uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32);
        const int rows = (int)((a.n_paths - path0) < 32 ? (a.n_paths - path0) : 32);
        // ---- chunk-local log returns of steps 32c .. 32c + 31 ----------------------------------------------------
        R xl[32];
        R x = DEG ? (R)1 : (R)0;
        auto growth = [&](R radw, R trig) {      // DEG > 0: x <- x exp(wc rad trig + dc), chunk-local running product
            const R d = fmaf((float)radw, (float)trig, (float)dc);
            R q = DEG >= 5 ? (R)(1.0 / 120.0) : (R)(1.0 / 24.0);
            if (DEG >= 5) q = fmaf(d, q, (R)(1.0 / 24.0));
            q = fmaf(d, q, (R)(1.0 / 6.0));
            q = fmaf(d, q, (R)0.5);
            q = fmaf(d, q, (R)1);
            x = fmaf(x, d * q, x);
        };
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const U4 u = philox4x32_10(c0, c1, (uint32_t)(4 * c + b), B200MC_STREAM_GBM, a.key);
            const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if constexpr (DEG > 0) {
                    // the step weight is folded into the radius once per pair (one FMUL instead of two products)
                    const BM3 bp = box_muller_parts(ww[t]);
                    const R radw = (R)bp.rad * wc;
                    growth(radw, (R)bp.cs);
                    xl[8 * b + 2 * t] = x;
                    growth(radw, (R)bp.sn);
                    xl[8 * b + 2 * t + 1] = x;
                } else {
                    const BM2 bm = box_muller_word(ww[t]);
                    R wa, wb, da, db;
                    if constexpr (TAB) {
                        const int s = 32 * c + 8 * b + 2 * t;
                        wa = w2[s]; wb = w2[s + 1]; da = d2[s]; db = d2[s + 1];
                    } else { wa = wb = wc; da = db = dc; }
                    x = (x + da) + wa * (R)bm.rc;
                    xl[8 * b + 2 * t] = x;
                    x = (x + db) + wb * (R)bm.rs;
                    xl[8 * b + 2 * t + 1] = x;
                }
            }
        }
        totals[c * 32 + lane] = x;
        // the TMA engine must have finished READING the tile of the previous group before anyone refills it
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        R off = DEG ? S0 : (R)0;
        for (int cc = 0; cc < c; ++cc) {
            if constexpr (DEG > 0) off *= totals[cc * 32 + lane]; else off += totals[cc * 32 + lane];
        }
        // ---- fill: y_{32c + 1 + j} = S0 exp(off + xl[j]) -----------------------------------------------------------
        O *row = tile + (size_t)lane * ncol;
        if (c == 0) row[0] = (O)a.m.S0;                                             // column 0 = S0   (:217)
        const int colbase = 32 * c + 1;
        if (colbase + 31 < ncol) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if constexpr (DEG > 0) row[colbase + j] = (O)(off * xl[j]);
                else if constexpr (sizeof(R) == 4) row[colbase + j] = (O)(S0 * ex2_approx((float)(off + xl[j])));
                else row[colbase + j] = (O)(S0 * exp(off + xl[j]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                O y;
                if constexpr (DEG > 0) y = (O)(off * xl[j]);
                else if constexpr (sizeof(R) == 4) y = (O)(S0 * ex2_approx((float)(off + xl[j])));
                else y = (O)(S0 * exp(off + xl[j]));
                if (colbase + j < ncol) row[colbase + j] = y;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // generic-proxy writes -> async proxy
        __syncthreads();
        // ---- flush ------------------------------------------------------------------------------------------------
        O *dst = out + (size_t)path0 * ncol;
        const uint32_t bytes = (uint32_t)((size_t)rows * ncol * sizeof(O));
        if (dst_aligned && (bytes & 15) == 0) {
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst), "r"(smem_u32(tile)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {                                                                    // ragged last group / odd base
            const int64_t E = (int64_t)rows * ncol;
            for (int64_t f = threadIdx.x; f < E; f += blockDim.x) dst[f] = tile[f];
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- host side
// Compile-time dispatch without macros: with_types / with_mode / with_bool call a generic lambda with tag values
// whose TYPES carry the template arguments (float{} / double{}, std::integral_constant, std::bool_constant).
template <typename F> static void with_types(bool fp64_state, int dtype, F &&f)
{
    if (fp64_state) { if (dtype == B200MC_F64) f(double{}, double{}); else f(double{}, float{}); }
    else { if (dtype == B200MC_F64) f(float{}, double{}); else f(float{}, float{}); }
}
template <typename F> static void with_mode(int mode, F &&f)
{
    switch (mode) {
    case MODE_GBM: f(std::integral_constant<int, MODE_GBM>{}); break;
    case MODE_DETVAR: f(std::integral_constant<int, MODE_DETVAR>{}); break;
    case MODE_HESTON: f(std::integral_constant<int, MODE_HESTON>{}); break;
    default: f(std::integral_constant<int, MODE_SVJ>{}); break;
    }
}
template <typename F> static void with_bool(bool b, F &&f)
{
    if (b) f(std::true_type{}); else f(std::false_type{});
}

template <typename K> static void allow_smem(K kernel, size_t smem)
{
    if (smem > 48 * 1024) cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

static int upload_tables(b200mc_handle *h, const Prep &pr, const double **wtab_d, const double **dtab_d)
{
    *wtab_d = nullptr;
    *dtab_d = nullptr;
    if (pr.wtab.empty()) return 0;
    const size_t nb = (pr.wtab.size() + pr.dtab.size()) * 8;
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, nb));
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, nb > 4096 ? nb : 4096, true));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(h->h_pinned, pr.wtab.data(), pr.wtab.size() * 8);
    memcpy((char *)h->h_pinned + pr.wtab.size() * 8, pr.dtab.data(), pr.dtab.size() * 8);
    B200MC_CUDA(h, cudaMemcpyAsync(h->d_scratch, h->h_pinned, nb, cudaMemcpyHostToDevice, h->stream));
    *wtab_d = (const double *)h->d_scratch;
    *dtab_d = *wtab_d + pr.wtab.size();
    return 0;
}

} // namespace b200mc
using namespace b200mc;

extern "C" int b200mc_simulate_terminal(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                        int32_t n_steps, int64_t n_paths, uint64_t seed, uint64_t path_offset,
                                        uint32_t flags, int dtype, int on_device, void *S_T, void *S_T_anti, void *v_T)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (dtype != B200MC_F32 && dtype != B200MC_F64) return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    if (flags & B200MC_GREEKS) return fail(h, B200MC_EINVAL, "B200MC_GREEKS is not valid for simulate_terminal");
    if (flags & B200MC_WIDE_RNG) return fail(h, B200MC_EINVAL, "B200MC_WIDE_RNG is valid for b200mc_price_european only");
    Prep pr;
    B200MC_TRY(prepare(h, p, S0, T, n_steps, n_paths, seed, flags, nullptr, pr));
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const bool anti = flags & B200MC_ANTITHETIC, fp64 = flags & B200MC_FP64;
    if (S_T_anti && !anti) return fail(h, B200MC_EINVAL, "S_T_anti given without B200MC_ANTITHETIC");
    const size_t esz = dtype == B200MC_F64 ? 8 : 4;
    const double *wtab_d, *dtab_d;
    B200MC_TRY(upload_tables(h, pr, &wtab_d, &dtab_d));
    PathArgs a;
    memset(&a, 0, sizeof(a));
    a.m = pr.m; a.key = pr.key; a.path0 = path_offset; a.n_paths = n_paths; a.n_steps = n_steps; a.wld = pr.wld;
    if (pr.mode == MODE_DETVAR) {   // final variance of the primary state, computed by the host recurrence
        double v = p->v0;
        for (int s = 0; s < n_steps; ++s) {
            const double vp = v > 0.0 ? v : 0.0;
            v = vp + p->kappa * (p->theta - vp) * pr.m.dt;
            v = v > 0.0 ? v : 0.0;
        }
        a.m.v0[1] = v;
    }
    void *dS = S_T, *dA = S_T_anti, *dV = v_T;
    if (!on_device) {
        const int nout = (S_T ? 1 : 0) + (S_T_anti ? 1 : 0) + (v_T ? 1 : 0);
        if (nout == 0) return fail(h, B200MC_EINVAL, "no output requested");
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)nout * n_paths * esz + 256));
        char *b = (char *)h->d_stage;
        if (S_T) { dS = b; b += (size_t)n_paths * esz; }
        if (S_T_anti) { dA = b; b += (size_t)n_paths * esz; }
        if (v_T) { dV = b; }
    }
    const size_t smem = pr.mode == MODE_DETVAR ? (size_t)3 * pr.wld * (fp64 ? 8 : 4) : 0;
    if (smem > 200 * 1024) return fail(h, B200MC_EINVAL, "too many steps for the deterministic-variance tables");
    int64_t grid = (n_paths + PT_THREADS - 1) / PT_THREADS;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (grid > cap) grid = cap;
    with_mode(pr.mode, [&](auto mode) {
        with_bool(anti, [&](auto twin) {
            with_types(fp64, dtype, [&](auto r, auto o) {
                using R = decltype(r);
                using O = decltype(o);
                auto k = k_terminal<decltype(mode)::value, decltype(twin)::value, R, O>;
                allow_smem(k, smem);
                k<<<(unsigned)grid, PT_THREADS, smem, h->stream>>>(a, wtab_d, (O *)dS, (O *)dA, (O *)dV);
            });
        });
    });
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (!on_device) {
        if (S_T) B200MC_CUDA(h, cudaMemcpyAsync(S_T, dS, (size_t)n_paths * esz, cudaMemcpyDeviceToHost, h->stream));
        if (S_T_anti) B200MC_CUDA(h, cudaMemcpyAsync(S_T_anti, dA, (size_t)n_paths * esz, cudaMemcpyDeviceToHost, h->stream));
        if (v_T) B200MC_CUDA(h, cudaMemcpyAsync(v_T, dV, (size_t)n_paths * esz, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" int b200mc_generate_paths(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                                     int64_t n_paths, uint64_t seed, uint64_t path_offset, uint32_t flags, int dtype,
                                     int on_device, void *out, int64_t ld)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (dtype != B200MC_F32 && dtype != B200MC_F64) return fail(h, B200MC_EINVAL, "dtype must be B200MC_F32 or B200MC_F64");
    if (flags & (B200MC_GREEKS | B200MC_ANTITHETIC | B200MC_WIDE_RNG))
        return fail(h, B200MC_EINVAL, "generate_paths takes only B200MC_FP64 / B200MC_FORCE_SVJ");
    if (!out) return fail(h, B200MC_EINVAL, "out is NULL");
    if (ld < (int64_t)n_steps + 1) return fail(h, B200MC_EINVAL, "ld must be at least n_steps + 1");
    Prep pr;
    B200MC_TRY(prepare(h, p, S0, T, n_steps, n_paths, seed, flags, nullptr, pr));
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const bool fp64 = flags & B200MC_FP64;
    const size_t esz = dtype == B200MC_F64 ? 8 : 4, rsz = fp64 ? 8 : 4;
    const double *wtab_d, *dtab_d;
    B200MC_TRY(upload_tables(h, pr, &wtab_d, &dtab_d));
    PathArgs a;
    memset(&a, 0, sizeof(a));
    a.m = pr.m; a.key = pr.key; a.path0 = path_offset; a.n_paths = n_paths; a.n_steps = n_steps; a.wld = pr.wld; a.ld = ld;
    void *dO = out;
    if (!on_device) {
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n_paths * ld * esz + 256));
        dO = h->d_stage;
    }
    constexpr int NW = PT_THREADS / 32;
    const int64_t groups = (n_paths + 31) / 32;                       // 32 paths per warp (or per CTA in the TMA kernel)
    const bool det = pr.mode == MODE_GBM || pr.mode == MODE_DETVAR;
    const bool tab = pr.mode == MODE_DETVAR;
    const size_t tab_bytes = tab ? (size_t)2 * (pr.wld + 32) * rsz : 0;

    if (det) {
        const int nch = (n_steps + 31) / 32;                          // TMA kernel: warps per CTA = chunks of 32 steps
        const size_t tma_smem = (((size_t)32 * (n_steps + 1) * esz + 127) & ~(size_t)127) + (size_t)nch * 32 * rsz + tab_bytes;
        const size_t row_smem = (size_t)NW * 32 * 33 * esz + tab_bytes;
        if (ld == (int64_t)n_steps + 1 && n_steps <= 1024 && tma_smem <= 200 * 1024) {
            // reference layout: CTA tile = 32 whole paths, one TMA bulk store per tile
            const int64_t grid = std::min<int64_t>(groups, (int64_t)h->sm_count * 16);
            // fp32 state, constant variance: carry the spot multiplicatively with a polynomial expm1 when every
            // |delta| = |drift dt| + sigma sqrt(dt) * 4.8 (largest raw draw) keeps the Taylor remainder below 2e-7 per step
            int deg = 0;
            if (!tab && !fp64) {
                const double dmax = fabs(pr.m.step_drift[0]) + fabs(pr.m.x_w[0]) * 4.8;
                const char *e = getenv("B200MC_PATHS_EXP");              // tuning knob: mufu | poly4 | poly5
                if (e && !strcmp(e, "mufu")) deg = 0;
                else if (e && !strcmp(e, "poly5")) deg = dmax <= 0.30 ? 5 : 0;
                else if (e && !strcmp(e, "poly4")) deg = dmax <= 0.30 ? 4 : 0;
                else deg = B200MC_PATHS_DEFAULT_POLY ? (dmax <= 0.125 ? 4 : (dmax <= 0.30 ? 5 : 0)) : 0;
            }
            const char *mb = getenv("B200MC_PATHS_MINB");               // tuning knob: resident CTAs per SM asked of ptxas
            const bool minb4 = !tab && !fp64 && nch <= 8 && (mb ? atoi(mb) == 4 : B200MC_PATHS_DEFAULT_MINB == 4);
            auto launch = [&](auto t, auto big, auto r, auto o, auto dg) {
                using O = decltype(o);
                using Rr = decltype(r);
                constexpr bool T_ = decltype(t)::value, B_ = decltype(big)::value;
                constexpr int D_ = decltype(dg)::value;
                if constexpr (!T_ && !B_ && sizeof(Rr) == 4) {
                    if (minb4) {
                        auto k = k_paths_tma<false, Rr, O, 256, D_, 4>;
                        allow_smem(k, tma_smem);
                        k<<<(unsigned)grid, 32 * nch, tma_smem, h->stream>>>(a, wtab_d, dtab_d, (O *)dO);
                        return;
                    }
                }
                auto k = k_paths_tma<T_, Rr, O, B_ ? 1024 : 256, D_>;
                allow_smem(k, tma_smem);
                k<<<(unsigned)grid, 32 * nch, tma_smem, h->stream>>>(a, wtab_d, dtab_d, (O *)dO);
            };
            with_bool(nch > 8, [&](auto big) {
                if (deg) {
                    with_bool(dtype == B200MC_F64, [&](auto wide) {
                        using O = std::conditional_t<decltype(wide)::value, double, float>;
                        if (deg == 5) launch(std::false_type{}, big, float{}, O{}, std::integral_constant<int, 5>{});
                        else launch(std::false_type{}, big, float{}, O{}, std::integral_constant<int, 4>{});
                    });
                } else {
                    with_bool(tab, [&](auto t) {
                        with_types(fp64, dtype, [&](auto r, auto o) { launch(t, big, r, o, std::integral_constant<int, 0>{}); });
                    });
                }
            });
        } else {
            // padded rows (or very long paths): row-tiled kernel, 128-bit stores when the rows allow them
            if (row_smem > 200 * 1024) return fail(h, B200MC_EINVAL, "too many steps for the deterministic-variance tables");
            const bool vec = dtype == B200MC_F32 && (ld % 4 == 0) && (((uintptr_t)dO) % 16 == 0);
            const int64_t grid = std::min<int64_t>((groups + NW - 1) / NW, (int64_t)h->sm_count * 8);
            with_bool(tab, [&](auto t) {
                with_types(fp64, dtype, [&](auto r, auto o) {
                    using O = decltype(o);
                    if constexpr (sizeof(O) == 4) {
                        if (vec) {
                            auto k = k_paths_det<decltype(t)::value, decltype(r), O, true>;
                            allow_smem(k, row_smem);
                            k<<<(unsigned)grid, PT_THREADS, row_smem, h->stream>>>(a, wtab_d, dtab_d, (O *)dO);
                            return;
                        }
                    }
                    auto k = k_paths_det<decltype(t)::value, decltype(r), O, false>;
                    allow_smem(k, row_smem);
                    k<<<(unsigned)grid, PT_THREADS, row_smem, h->stream>>>(a, wtab_d, dtab_d, (O *)dO);
                });
            });
        }
    } else {
        // Heston / SVJ: one lane per path, per-row rings of two aligned 64-byte windows
        const size_t smem = (size_t)NW * 32 * (2 * (64 / esz) + 1) * esz;
        const int64_t grid = std::min<int64_t>((groups + NW - 1) / NW, (int64_t)h->sm_count * 8);
        with_bool(pr.mode == MODE_SVJ, [&](auto jumps) {
            with_types(fp64, dtype, [&](auto r, auto o) {
                using O = decltype(o);
                auto k = k_paths<decltype(jumps)::value ? MODE_SVJ : MODE_HESTON, decltype(r), O>;
                allow_smem(k, smem);
                k<<<(unsigned)grid, PT_THREADS, smem, h->stream>>>(a, wtab_d, dtab_d, (O *)dO);
            });
        });
    }
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (!on_device) {
        B200MC_CUDA(h, cudaMemcpyAsync(out, dO, (size_t)n_paths * ld * esz, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}
