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
#include "prep.cuh"

namespace b200mc {

constexpr int PT_THREADS = 256;
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

// ---------------------------------------------------------------------------------------------- path store
template <int MODE, typename R, typename O> struct TileRec {
    static constexpr bool enabled = true;
    O *tile;             // this warp's [32][33] tile
    O *out;              // global matrix
    const R *dtab;       // DETVAR: cumulative drift
    int64_t path0, n_paths, ld;
    int n_steps, lane;
    R S0;

    __device__ __noinline__ void flush(int s) const
    {
        const int c0 = s & ~31, nc = (s & 31) + 1;
        __syncwarp();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const int64_t p = path0 + r;
            if (p < n_paths && lane < nc) out[(size_t)p * ld + 1 + c0 + lane] = tile[r * 33 + lane];
        }
        __syncwarp();
    }
    __device__ __forceinline__ void operator()(int s, R x) const
    {
        if constexpr (MODE == MODE_DETVAR) x += dtab[s];
        R S;
        if constexpr (sizeof(R) == 4) S = S0 * exp2f(x * LOG2E_F);
        else S = S0 * exp(x);
        tile[lane * 33 + (s & 31)] = (O)S;
        if ((s & 31) == 31 || s == n_steps - 1) flush(s);
    }
};

template <int MODE, typename R, typename O>
__global__ void __launch_bounds__(PT_THREADS)
k_paths(const __grid_constant__ PathArgs a, const double *__restrict__ wtab_g, const double *__restrict__ dtab_g,
        O *__restrict__ out)
{
    using L = StateLayout<false, false>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    O *tiles = reinterpret_cast<O *>(smem_raw);                                 // [8][32*33]
    R *wtab = reinterpret_cast<R *>(tiles + (PT_THREADS / 32) * 32 * 33);       // [3][wld]
    R *dtab = wtab + 3 * a.wld;                                                 // [wld]
    if constexpr (MODE == MODE_DETVAR) {
        for (int i = threadIdx.x; i < 3 * a.wld; i += PT_THREADS) wtab[i] = (R)wtab_g[i];
        for (int i = threadIdx.x; i < a.wld; i += PT_THREADS) dtab[i] = (R)dtab_g[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_groups = (a.n_paths + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * (PT_THREADS / 32) + warp; g < n_groups;
         g += (int64_t)gridDim.x * (PT_THREADS / 32)) {
        TileRec<MODE, R, O> rec;
        rec.tile = tiles + warp * 32 * 33;
        rec.out = out;
        rec.dtab = dtab;
        rec.path0 = g * 32;
        rec.n_paths = a.n_paths;
        rec.ld = a.ld;
        rec.n_steps = a.n_steps;
        rec.lane = lane;
        rec.S0 = (R)a.m.S0;
        int64_t me = rec.path0 + lane;
        if (me < a.n_paths) out[(size_t)me * a.ld] = (O)a.m.S0;                 // column 0 = S0   (:217)
        else me = a.n_paths - 1;   // idle lanes shadow the last path so the warp stays convergent in flush()
        R xT[L::NS], vT[L::NS], sumz;
        simulate_path<MODE, false, false, R>(a.m, a.key, a.path0 + (uint64_t)me, a.n_steps, wtab, a.wld, xT, vT, sumz,
                                             rec);
    }
}

using TermKernel = void (*)(const PathArgs, const double *, void *, void *, void *);

template <int MODE, bool ANTI, typename R, typename O>
static void term_launch(const PathArgs &a, const double *wtab, void *s, void *sa, void *v, unsigned grid, size_t smem,
                        cudaStream_t st)
{
    auto k = k_terminal<MODE, ANTI, R, O>;
    if (smem > 48 * 1024) cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, PT_THREADS, smem, st>>>(a, wtab, (O *)s, (O *)sa, (O *)v);
}
template <int MODE, typename R, typename O>
static void path_launch(const PathArgs &a, const double *wtab, const double *dtab, void *out, unsigned grid, size_t smem,
                        cudaStream_t st)
{
    auto k = k_paths<MODE, R, O>;
    cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, PT_THREADS, smem, st>>>(a, wtab, dtab, (O *)out);
}

#define DISPATCH_MODE(mode, CALL)                       \
    switch (mode) {                                     \
    case MODE_GBM: { constexpr int M = MODE_GBM; CALL; } break;       \
    case MODE_DETVAR: { constexpr int M = MODE_DETVAR; CALL; } break; \
    case MODE_HESTON: { constexpr int M = MODE_HESTON; CALL; } break; \
    default: { constexpr int M = MODE_SVJ; CALL; } break;             \
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
    int64_t grid = (n_paths + PT_THREADS - 1) / PT_THREADS;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (grid > cap) grid = cap;
#define TERM_CALL                                                                                            \
    do {                                                                                                     \
        if (fp64) {                                                                                          \
            if (dtype == B200MC_F64) { if (anti) term_launch<M, true, double, double>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); else term_launch<M, false, double, double>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); } \
            else { if (anti) term_launch<M, true, double, float>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); else term_launch<M, false, double, float>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); } \
        } else {                                                                                             \
            if (dtype == B200MC_F64) { if (anti) term_launch<M, true, float, double>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); else term_launch<M, false, float, double>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); } \
            else { if (anti) term_launch<M, true, float, float>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); else term_launch<M, false, float, float>(a, wtab_d, dS, dA, dV, (unsigned)grid, smem, h->stream); } \
        }                                                                                                    \
    } while (0)
    DISPATCH_MODE(pr.mode, TERM_CALL);
#undef TERM_CALL
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
    if (flags & (B200MC_GREEKS | B200MC_ANTITHETIC))
        return fail(h, B200MC_EINVAL, "generate_paths takes only B200MC_FP64 / B200MC_FORCE_SVJ");
    if (!out) return fail(h, B200MC_EINVAL, "out is NULL");
    if (ld < (int64_t)n_steps + 1) return fail(h, B200MC_EINVAL, "ld must be at least n_steps + 1");
    Prep pr;
    B200MC_TRY(prepare(h, p, S0, T, n_steps, n_paths, seed, flags, nullptr, pr));
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const bool fp64 = flags & B200MC_FP64;
    const size_t esz = dtype == B200MC_F64 ? 8 : 4;
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
    size_t smem = (size_t)(PT_THREADS / 32) * 32 * 33 * esz;
    if (pr.mode == MODE_DETVAR) smem += (size_t)4 * pr.wld * (fp64 ? 8 : 4);
    if (smem > 200 * 1024) return fail(h, B200MC_EINVAL, "too many steps for the deterministic-variance tables");
    int64_t grid = ((n_paths + 31) / 32 + PT_THREADS / 32 - 1) / (PT_THREADS / 32);
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (grid > cap) grid = cap;
#define PATH_CALL                                                                                              \
    do {                                                                                                       \
        if (fp64) { if (dtype == B200MC_F64) path_launch<M, double, double>(a, wtab_d, dtab_d, dO, (unsigned)grid, smem, h->stream); \
                    else path_launch<M, double, float>(a, wtab_d, dtab_d, dO, (unsigned)grid, smem, h->stream); }  \
        else { if (dtype == B200MC_F64) path_launch<M, float, double>(a, wtab_d, dtab_d, dO, (unsigned)grid, smem, h->stream); \
               else path_launch<M, float, float>(a, wtab_d, dtab_d, dO, (unsigned)grid, smem, h->stream); }         \
    } while (0)
    DISPATCH_MODE(pr.mode, PATH_CALL);
#undef PATH_CALL
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    if (!on_device) {
        B200MC_CUDA(h, cudaMemcpyAsync(out, dO, (size_t)n_paths * ld * esz, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}
