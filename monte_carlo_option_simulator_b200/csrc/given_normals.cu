// given_normals.cu -- deterministic mode: the SVJ recurrence on caller-supplied draws.
//
// Drop-in for _simulate_svj_paths_numba (engine/monte_carlo.py:189-243): same inputs (four C-contiguous float64
// [n_paths, n_steps] arrays), same outputs, fp64, and the reference's own operation order -- every product and
// sum below is a separately rounded IEEE operation (__dmul_rn / __dadd_rn, no FMA contraction), S is updated
// multiplicatively with one exp per step (:236).  Per-path results therefore differ from the reference only by
// the last-bit behaviour of exp().
//
// The arrays are path-major, so "one thread per path" would read with a stride of n_steps*8 bytes.  Each warp
// instead owns 32 consecutive paths and walks the time axis in tiles of 8 steps: a quarter-warp loads one path's
// 8 consecutive doubles (two full sectors), the tile is parked in shared memory with a padded pitch, and each lane
// then reads its own row.  The optional path record goes the other way through the same kind of tile.
// Arrays the parameters make irrelevant (Z2 when xi == 0, the jump arrays when lambda_j <= 0) are never read and get
// no tile: shared memory per warp is 2.3 KB per array in use, which is what sets the number of resident warps -- the
// per-step fp64 chain (sqrt, exp, separately rounded products) has a latency of several hundred cycles and needs them.
#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200mc {

constexpr int GN_WARPS = 4;

struct GivenArgs {
    double S0, v0, dt, sqrt_dt, drift_comp, kappa, theta, xi, rho, sq1mr2, jump_thr, mu_j, sigma_j;
    double zsign;            // +1, or -1 for the antithetic twin: -Z1, -Z2, U, -Z_jump_size (monte_carlo.py:318-324); exact
    int64_t n_paths;
    int32_t n_steps;
    int32_t need_z2, need_jump, record;
};

// lane (sub = lane >> 3, col = lane & 7) fetches element [path0 + 4 it + sub][s0 + col] for it = 0..7.  `rowptr` already
// points at [path0 + sub][col] of the array, so the address of each load is one 64-bit add away (the full index
// arithmetic per load costs more issue slots than the fp64 step itself).
template <int TS>
__device__ __forceinline__ void load_tile(double *tile, const double *__restrict__ rowptr, size_t stride_r, int s0,
                                          bool full, int rows_left, int cols_left, int lane)
{
    constexpr int RPI = 32 / TS, PITCH = TS + 1;            // rows per load instruction
    const int col = lane % TS, sub = lane / TS;
    const double *p = rowptr + s0;
    if (full) {
#pragma unroll
        for (int it = 0; it < TS; ++it) {
            tile[(RPI * it + sub) * PITCH + col] = __ldg(p);
            p += stride_r;
        }
    } else {
#pragma unroll
        for (int it = 0; it < TS; ++it) {
            const int row = RPI * it + sub;
            tile[row * PITCH + col] = (row < rows_left && col < cols_left) ? __ldg(p) : 0.0;
            p += stride_r;
        }
    }
}

// ILP: a full tile is walked in three unrolled sweeps -- the variance recurrence and the exponent of every step (the
// only chain that runs from step to step besides one multiply), then the GN_TS exponentials (independent of each other:
// the fp64 pipe sees GN_TS chains per lane instead of one), then the running product.  The same separately rounded
// operations per path in the same order as the step-by-step loop, which partial tiles still take: results are bitwise
// the same.
#ifndef GN_ILP_MINB8
#define GN_ILP_MINB8 6
#endif
template <int GN_TS, bool ILP>
__global__ void __launch_bounds__(GN_WARPS * 32, GN_TS == 8 ? (ILP ? GN_ILP_MINB8 : 6) : 3)
k_given_normals(const __grid_constant__ GivenArgs a, const double *__restrict__ Z1, const double *__restrict__ Z2,
                const double *__restrict__ Zj, const double *__restrict__ Zjs, double *__restrict__ S_final,
                double *__restrict__ v_final, double *__restrict__ all_paths)
{
    extern __shared__ __align__(16) double gn_tiles[];          // [GN_WARPS][tiles in use][32 * GN_PITCH]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GN_PITCH = GN_TS + 1, TILE = 32 * GN_PITCH;
    const int ntile = 1 + a.need_z2 + 2 * a.need_jump + a.record;
    double *t1 = gn_tiles + (size_t)warp * ntile * TILE;
    double *t2 = t1 + TILE;                                   // valid only when need_z2
    double *tj = t1 + (1 + a.need_z2) * TILE, *tjs = tj + TILE;   // valid only when need_jump
    double *tout = t1 + (ntile - 1) * TILE;                   // valid only when record
    const int64_t n_groups = (a.n_paths + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * GN_WARPS + warp; g < n_groups; g += (int64_t)gridDim.x * GN_WARPS) {
        const int64_t path0 = g * 32, me = path0 + lane;
        double S = a.S0, v = a.v0;                                              // :212-213
        if (a.record) {
            if (me < a.n_paths) all_paths[(size_t)me * (a.n_steps + 1)] = a.S0;  // :217
        }
        const int rows_left = (int)((a.n_paths - path0) < 32 ? (a.n_paths - path0) : 32);
        const size_t lane_off = (size_t)(path0 + lane / GN_TS) * a.n_steps + (lane % GN_TS),
                     stride4 = (size_t)(32 / GN_TS) * a.n_steps;
        for (int s0 = 0; s0 < a.n_steps; s0 += GN_TS) {
            const int cols_left = a.n_steps - s0;
            const bool full = rows_left == 32 && cols_left >= GN_TS;
            __syncwarp();
            load_tile<GN_TS>(t1, Z1 + lane_off, stride4, s0, full, rows_left, cols_left, lane);
            if (a.need_z2) load_tile<GN_TS>(t2, Z2 + lane_off, stride4, s0, full, rows_left, cols_left, lane);
            if (a.need_jump) {
                load_tile<GN_TS>(tj, Zj + lane_off, stride4, s0, full, rows_left, cols_left, lane);
                load_tile<GN_TS>(tjs, Zjs + lane_off, stride4, s0, full, rows_left, cols_left, lane);
            }
            __syncwarp();
            const int ns = min(GN_TS, a.n_steps - s0);
            if (ILP && ns == GN_TS) {
                double e[GN_TS];
#pragma unroll
                for (int t = 0; t < GN_TS; ++t) {
                    const double z1 = __dmul_rn(a.zsign, t1[lane * GN_PITCH + t]);
                    const double v_pos = fmax(v, 0.0);                           // :223
                    const double sqrt_v = sqrt(v_pos);                           // :224
                    const double dW1 = __dmul_rn(z1, a.sqrt_dt);                 // :226
                    double dW2 = 0.0;                                            // :227
                    if (a.need_z2)
                        dW2 = __dadd_rn(__dmul_rn(__dmul_rn(a.rho, z1), a.sqrt_dt),
                                        __dmul_rn(__dmul_rn(a.sq1mr2, __dmul_rn(a.zsign, t2[lane * GN_PITCH + t])), a.sqrt_dt));
                    const double log_drift = __dmul_rn(__dadd_rn(a.drift_comp, -__dmul_rn(0.5, v_pos)), a.dt);   // :229
                    const double log_diff = __dmul_rn(sqrt_v, dW1);              // :230
                    double jump = 0.0;                                           // :232
                    if (a.need_jump) {
                        if (tj[lane * GN_PITCH + t] < a.jump_thr)                // :233
                            jump = __dadd_rn(a.mu_j, __dmul_rn(a.sigma_j, __dmul_rn(a.zsign, tjs[lane * GN_PITCH + t])));
                    }
                    e[t] = __dadd_rn(__dadd_rn(log_drift, log_diff), jump);
                    const double mr = __dmul_rn(__dmul_rn(a.kappa, __dadd_rn(a.theta, -v_pos)), a.dt);
                    const double vv = __dmul_rn(__dmul_rn(a.xi, sqrt_v), dW2);
                    v = fmax(__dadd_rn(__dadd_rn(v_pos, mr), vv), 0.0);          // :237-238
                }
#pragma unroll
                for (int t = 0; t < GN_TS; ++t) e[t] = exp(e[t]);
#pragma unroll
                for (int t = 0; t < GN_TS; ++t) {
                    S = __dmul_rn(S, e[t]);                                      // :236
                    if (a.record) tout[lane * GN_PITCH + t] = S;
                }
            } else
            for (int t = 0; t < ns; ++t) {
                const double z1 = __dmul_rn(a.zsign, t1[lane * GN_PITCH + t]);
                const double v_pos = fmax(v, 0.0);                               // :223
                const double sqrt_v = sqrt(v_pos);                               // :224
                const double dW1 = __dmul_rn(z1, a.sqrt_dt);                     // :226
                double dW2 = 0.0;                                                // :227 (unused when xi == 0)
                if (a.need_z2)
                    dW2 = __dadd_rn(__dmul_rn(__dmul_rn(a.rho, z1), a.sqrt_dt),
                                    __dmul_rn(__dmul_rn(a.sq1mr2, __dmul_rn(a.zsign, t2[lane * GN_PITCH + t])), a.sqrt_dt));
                const double log_drift = __dmul_rn(__dadd_rn(a.drift_comp, -__dmul_rn(0.5, v_pos)), a.dt);   // :229
                const double log_diff = __dmul_rn(sqrt_v, dW1);                  // :230
                double jump = 0.0;                                               // :232
                if (a.need_jump) {
                    if (tj[lane * GN_PITCH + t] < a.jump_thr)                    // :233
                        jump = __dadd_rn(a.mu_j, __dmul_rn(a.sigma_j, __dmul_rn(a.zsign, tjs[lane * GN_PITCH + t])));   // :234
                }
                S = __dmul_rn(S, exp(__dadd_rn(__dadd_rn(log_drift, log_diff), jump)));              // :236
                const double mr = __dmul_rn(__dmul_rn(a.kappa, __dadd_rn(a.theta, -v_pos)), a.dt);
                const double vv = __dmul_rn(__dmul_rn(a.xi, sqrt_v), dW2);
                v = fmax(__dadd_rn(__dadd_rn(v_pos, mr), vv), 0.0);              // :237-238
                if (a.record) tout[lane * GN_PITCH + t] = S;
            }
            if (a.record) {                                                      // :241, coalesced by rows
                __syncwarp();
                const int col = lane % GN_TS, sub = lane / GN_TS;
#pragma unroll
                for (int it = 0; it < GN_TS; ++it) {
                    const int row = (32 / GN_TS) * it + sub;
                    const int64_t p = path0 + row;
                    if (p < a.n_paths && col < ns)
                        all_paths[(size_t)p * (a.n_steps + 1) + 1 + s0 + col] = tout[row * GN_PITCH + col];
                }
            }
        }
        if (me < a.n_paths) {
            S_final[me] = S;
            v_final[me] = v;
        }
    }
}

// Host -> device copy of a large PAGEABLE array at PCIe speed.  A plain cudaMemcpy from pageable memory is staged by
// the driver through one thread (measured 11 GB/s); here H2D_THREADS workers each own two pinned 4 MiB buffers and a
// stream: they memcpy their slices into pinned memory and hand them to the DMA engine, so the CPU copy is parallel and
// overlaps the transfers.  The pinned pool (64 MiB) is allocated once per handle.
constexpr int H2D_THREADS = 8;
constexpr size_t H2D_CHUNK = (size_t)4 << 20;

static int parallel_h2d(b200mc_handle *h, void *dst, const void *src, size_t bytes)
{
    // the destination (d_stage) may still be read by work queued on the handle's stream; the workers below copy on
    // private streams, so that work has to drain first
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    if (bytes < 8 * H2D_CHUNK) {
        B200MC_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    B200MC_TRY(ensure(h, &h->h_pool, &h->h_pool_bytes, (size_t)H2D_THREADS * 2 * H2D_CHUNK, true));
    const size_t n_chunks = (bytes + H2D_CHUNK - 1) / H2D_CHUNK;
    std::vector<cudaError_t> err(H2D_THREADS, cudaSuccess);
    std::vector<std::thread> workers;
    const int device = h->device;
    char *pool = (char *)h->h_pool;
    for (int t = 0; t < H2D_THREADS; ++t) {
        workers.emplace_back([=, &err]() {
            cudaError_t e = cudaSetDevice(device);
            cudaStream_t st = nullptr;
            cudaEvent_t ev[2] = {nullptr, nullptr};
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
            char *buf[2] = {pool + (size_t)(2 * t) * H2D_CHUNK, pool + (size_t)(2 * t + 1) * H2D_CHUNK};
            int k = 0;
            for (size_t c = t; c < n_chunks && e == cudaSuccess; c += H2D_THREADS, k ^= 1) {
                const size_t off = c * H2D_CHUNK, len = std::min(H2D_CHUNK, bytes - off);
                e = cudaEventSynchronize(ev[k]);                       // the DMA that last read this buffer is done
                if (e != cudaSuccess) break;
                memcpy(buf[k], (const char *)src + off, len);
                e = cudaMemcpyAsync((char *)dst + off, buf[k], len, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(ev[k], st);
            }
            if (st) {
                const cudaError_t e2 = cudaStreamSynchronize(st);
                if (e == cudaSuccess) e = e2;
                cudaStreamDestroy(st);
            }
            for (int q = 0; q < 2; ++q)
                if (ev[q]) cudaEventDestroy(ev[q]);
            err[t] = e;
        });
    }
    for (auto &w : workers) w.join();
    for (int t = 0; t < H2D_THREADS; ++t)
        if (err[t] != cudaSuccess) return fail(h, B200MC_ECUDA, "parallel host->device copy failed: %s", cudaGetErrorString(err[t]));
    return 0;
}

static int check_given(b200mc_handle *h, const b200mc_svj_params *p, int64_t n_paths, int32_t n_steps, double T,
                       const double *Z1, const double *Z2, const double *Zj, const double *Zjs, int record,
                       double *S_final, double *v_final, double *all_paths)
{
    if (!p) return fail(h, B200MC_EINVAL, "params is NULL");
    if (n_paths < 0 || n_steps <= 0) return fail(h, B200MC_EINVAL, "n_paths must be >= 0 and n_steps > 0");
    if (!(T == T)) return fail(h, B200MC_EINVAL, "T is NaN");
    if (n_paths > 0 && (!Z1 || !Z2 || !Zj || !Zjs || !S_final || !v_final))
        return fail(h, B200MC_EINVAL, "NULL array argument");
    if ((record & B200MC_GIVEN_RECORD) && n_paths > 0 && !all_paths) return fail(h, B200MC_EINVAL, "record_paths set but all_paths is NULL");
    if (record & ~(B200MC_GIVEN_RECORD | B200MC_GIVEN_NEGATE)) return fail(h, B200MC_EINVAL, "record_paths: unknown flag bits");
    return 0;
}

static GivenArgs make_args(const b200mc_svj_params *p, double S0, double T, int64_t n_paths, int32_t n_steps, int record)
{
    GivenArgs a;
    a.S0 = S0;
    a.v0 = p->v0;
    a.dt = T / (double)n_steps;                                              // :206
    a.sqrt_dt = sqrt(a.dt);                                                  // :207
    const double k = exp(p->mu_j + 0.5 * (p->sigma_j * p->sigma_j)) - 1.0;   // :209
    a.drift_comp = p->r - p->q - p->lambda_j * k;                            // :210
    a.kappa = p->kappa; a.theta = p->theta; a.xi = p->xi; a.rho = p->rho;
    a.sq1mr2 = sqrt(1.0 - p->rho * p->rho);
    a.jump_thr = p->lambda_j * a.dt;
    a.mu_j = p->mu_j; a.sigma_j = p->sigma_j;
    a.n_paths = n_paths; a.n_steps = n_steps;
    a.need_z2 = (p->xi != 0.0) ? 1 : 0;
    // Z_jump is a uniform in [0, 1): with lambda_j dt <= 0 the test :233 can never fire
    a.need_jump = (a.jump_thr > 0.0) ? 1 : 0;
    a.record = (record & B200MC_GIVEN_RECORD) ? 1 : 0;
    a.zsign = (record & B200MC_GIVEN_NEGATE) ? -1.0 : 1.0;
    return a;
}

static int launch_given(b200mc_handle *h, const GivenArgs &a, const double *Z1, const double *Z2, const double *Zj,
                        const double *Zjs, double *S_final, double *v_final, double *all_paths)
{
    if (a.n_paths == 0) return 0;
    const int64_t groups = (a.n_paths + 31) / 32;
    int64_t grid = (groups + GN_WARPS - 1) / GN_WARPS;
    const int64_t cap = (int64_t)h->sm_count * 16;
    if (grid > cap) grid = cap;
    const int ntile = 1 + a.need_z2 + 2 * a.need_jump + a.record;
    // 8-step tiles (64-byte row segments, 2.3 KB of shared memory per warp and array => more resident warps) unless the
    // run is memory heavy (jump arrays in use: 32 B per path-step) AND the rows are not 64-byte multiples: then 128-byte
    // segments waste fewer sectors (measured, 1M x 250 SVJ: 2.0 -> 2.5 TB/s; 4M x 64: 3.4 vs 2.9 TB/s the other way)
    const char *force = getenv("B200MC_GN_TILE");
    const bool wide = force ? atoi(force) == 16 : (a.need_jump != 0 && (a.n_steps % 8) != 0);
    const int ts = wide ? 16 : 8;
    const size_t smem = (size_t)GN_WARPS * ntile * 32 * (ts + 1) * sizeof(double);
    // three-sweep tile walk where the fp64 chain binds (one or two arrays: 1.18 -> 1.04 ms GBM, 1.41 -> 1.33 ms Heston at
    // 2.5e8 path-steps); with the jump arrays the kernel waits on memory and the sweeps cost 2-5 %
    // (profiles/r02_given_ilp_probe.txt).  B200MC_GN_ILP=0/1 forces either.
    const char *ie = getenv("B200MC_GN_ILP");
    const bool ilp = ie ? atoi(ie) != 0 : (a.need_jump == 0);
    auto kern = ilp ? (wide ? k_given_normals<16, true> : k_given_normals<8, true>)
                    : (wide ? k_given_normals<16, false> : k_given_normals<8, false>);
    B200MC_CUDA(h, cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, GN_WARPS * 32, smem, h->stream>>>(a, Z1, Z2, Zj, Zjs, S_final, v_final, all_paths);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

} // namespace b200mc
using namespace b200mc;

extern "C" int b200mc_simulate_given_normals_dev(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                                 int64_t n_paths, int32_t n_steps, const double *Z1, const double *Z2,
                                                 const double *Z_jump, const double *Z_jump_size, int record_paths,
                                                 double *S_final, double *v_final, double *all_paths)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_TRY(check_given(h, p, n_paths, n_steps, T, Z1, Z2, Z_jump, Z_jump_size, record_paths, S_final, v_final,
                           all_paths));
    B200MC_CUDA(h, cudaSetDevice(h->device));
    return launch_given(h, make_args(p, S0, T, n_paths, n_steps, record_paths), Z1, Z2, Z_jump, Z_jump_size, S_final,
                        v_final, all_paths);
}

extern "C" int b200mc_simulate_given_normals(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                             int64_t n_paths, int32_t n_steps, const double *Z1, const double *Z2,
                                             const double *Z_jump, const double *Z_jump_size, int record_paths,
                                             double *S_final, double *v_final, double *all_paths)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_TRY(check_given(h, p, n_paths, n_steps, T, Z1, Z2, Z_jump, Z_jump_size, record_paths, S_final, v_final,
                           all_paths));
    if (n_paths == 0) return 0;
    B200MC_CUDA(h, cudaSetDevice(h->device));
    GivenArgs a = make_args(p, S0, T, n_paths, n_steps, record_paths);
    const bool rec = (record_paths & B200MC_GIVEN_RECORD) != 0;
    const int narr = 1 + a.need_z2 + 2 * a.need_jump;
    // stage at most ~2 GiB of inputs + outputs per chunk of paths
    const size_t in_row = (size_t)n_steps * 8, out_row = 16 + (rec ? (size_t)(n_steps + 1) * 8 : 0);
    const size_t per_path = (size_t)narr * in_row + out_row;
    int64_t chunk = (int64_t)(((size_t)2 << 30) / per_path);
    chunk = chunk < 32 ? 32 : (chunk & ~(int64_t)31);
    if (chunk > n_paths) chunk = n_paths;
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)chunk * per_path + 1024));
    char *base = (char *)h->d_stage;
    double *dZ[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t off = 0;
    const int use[4] = {1, a.need_z2, a.need_jump, a.need_jump};
    for (int i = 0; i < 4; ++i) {
        if (use[i]) { dZ[i] = (double *)(base + off); off += (size_t)chunk * in_row; }
    }
    double *dS = (double *)(base + off); off += (size_t)chunk * 8;
    double *dv = (double *)(base + off); off += (size_t)chunk * 8;
    double *dP = rec ? (double *)(base + off) : nullptr;
    const double *src[4] = {Z1, Z2, Z_jump, Z_jump_size};
    for (int64_t p0 = 0; p0 < n_paths; p0 += chunk) {
        const int64_t np = (n_paths - p0 < chunk) ? (n_paths - p0) : chunk;
        for (int i = 0; i < 4; ++i) {
            if (use[i]) B200MC_TRY(parallel_h2d(h, dZ[i], src[i] + (size_t)p0 * n_steps, (size_t)np * in_row));
        }
        a.n_paths = np;
        B200MC_TRY(launch_given(h, a, dZ[0], dZ[1] ? dZ[1] : dZ[0], dZ[2] ? dZ[2] : dZ[0], dZ[3] ? dZ[3] : dZ[0], dS, dv,
                                dP));
        B200MC_CUDA(h, cudaMemcpyAsync(S_final + p0, dS, (size_t)np * 8, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaMemcpyAsync(v_final + p0, dv, (size_t)np * 8, cudaMemcpyDeviceToHost, h->stream));
        if (rec)
            B200MC_CUDA(h, cudaMemcpyAsync(all_paths + (size_t)p0 * (n_steps + 1), dP, (size_t)np * (n_steps + 1) * 8,
                                           cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}
