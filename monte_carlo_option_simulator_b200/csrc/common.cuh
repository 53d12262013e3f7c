// common.cuh -- handle, error plumbing and reduction helpers shared by the kernels of libb200mc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200mc.h"
#include "philox.cuh"

struct b200mc_handle {
    int device;
    int sm_count;
    int sm_clock_khz;
    int cc;
    uint64_t hbm_bytes;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int64_t launches;
    // scratch (grown on demand, never shrunk)
    void *d_scratch;        size_t scratch_bytes;      // block partials, weights, device-side results
    void *d_stage;          size_t stage_bytes;        // staging for host<->device array arguments
    void *h_pinned;         size_t pinned_bytes;       // pinned bounce buffer for small arguments
    void *d_result;         size_t result_bytes;       // device-side results of the synchronous entry points
    void *h_result;         size_t h_result_bytes;     // pinned landing buffer of those results
    void *h_pool;           size_t h_pool_bytes;       // pinned chunk pool of the parallel host->device copy
    bool own_stream;                                   // false after b200mc_set_stream
    int smem_optin;                                    // cudaDevAttrMaxSharedMemoryPerBlockOptin
    const void *occ_kern[64]; size_t occ_smem[64]; int occ_val[64]; int n_occ;   // (kernel, smem) -> resident CTAs per SM
    void *risk_state; bool risk_state_clean; unsigned long long risk_barrier;            // state of the one-launch tail-metric select (risk.cu), zeroed by the kernel itself
    const void *risk_x; int64_t risk_n; int risk_dtype; // vector of the multi-rank tail-metric primitives (risk.cu)
    unsigned int *d_counter;                           // "last block reduces" ticket
    void *peer_local; void *peer_ptr[B200MC_PEER_MAX_RANKS]; int peer_rank, peer_world;   // peer.cu: exchange buffers
    unsigned long long peer_epoch;
    void *peer_status, *peer_status_dev;               // pinned + mapped: epoch of the first exchange that timed out (0 = none)
    char err[512];
};

namespace b200mc {

extern thread_local char g_create_err[512];

inline int fail(b200mc_handle *h, int code, const char *fmt, const char *a = "", const char *b = "")
{
    char *dst = h ? h->err : g_create_err;
    snprintf(dst, 512, fmt, a, b);
    return code;
}

#define B200MC_CUDA(h, call)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? B200MC_ENOMEM : B200MC_ECUDA;    \
            return b200mc::fail((h), code__, "%s failed: %s", #call, cudaGetErrorString(e__)); \
        }                                                                                      \
    } while (0)

#define B200MC_TRY(expr)            \
    do {                            \
        int rc__ = (expr);          \
        if (rc__ != 0) return rc__; \
    } while (0)

inline int ensure(b200mc_handle *h, void **p, size_t *have, size_t want, bool pinned = false)
{
    if (*have >= want) return 0;
    if (*p) {
        // work queued on the handle's stream may still read the old buffer (an asynchronous copy out of the pinned bounce
        // buffer, a kernel on the scratch): drain it before the buffer goes away
        if (h && h->stream) cudaStreamSynchronize(h->stream);
        // the multi-rank tail-metric primitives keep their vector in d_stage and its keys in d_scratch between calls: once
        // either buffer is replaced, b200mc_risk_hist / _finish must fail loudly instead of reading freed memory
        if (h && (p == &h->d_stage || p == &h->d_scratch)) h->risk_x = nullptr;
        if (pinned) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    size_t sz = want + want / 4 + 4096;
    cudaError_t e = pinned ? cudaMallocHost(p, sz) : cudaMalloc(p, sz);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, B200MC_ENOMEM, "%s of scratch failed: %s", pinned ? "cudaMallocHost" : "cudaMalloc",
                    cudaGetErrorString(e));
    }
    *have = sz;
    return 0;
}

// Where the synchronous entry points (b200mc_price_european, b200mc_price_cells with host results) let the kernels write
// their few result doubles: "mapped" = straight into the pinned landing buffer h_result through the bus (pinned memory is
// device-addressable under unified addressing; the last CTA's 17 stores per strike are posted writes), so the call is
// launch -> synchronise -> memcpy with no device-to-host copy command in between; "copy" = into d_result, followed by a
// cudaMemcpyAsync.  Small calls (calibration: 1e4-1e5 of them, 10-40 us of GPU work each) are bound by this fixed cost.
// B200MC_RESULT=copy|mapped overrides the default; results larger than RESULT_MAPPED_MAX always take the copy.
#ifndef B200MC_RESULT_MAPPED_DEFAULT
#define B200MC_RESULT_MAPPED_DEFAULT 1
#endif
constexpr size_t RESULT_MAPPED_MAX = 64 * 1024;
inline bool result_mapped_wanted()
{
    static const int v = [] {
        const char *e = getenv("B200MC_RESULT");
        if (e && !strcmp(e, "mapped")) return 1;
        if (e && !strcmp(e, "copy")) return 0;
        return B200MC_RESULT_MAPPED_DEFAULT;
    }();
    return v != 0;
}
// device address of the pinned landing buffer, or nullptr when the results have to go through d_result
inline double *result_mapped_ptr(b200mc_handle *h, size_t bytes)
{
    if (!result_mapped_wanted() || bytes > RESULT_MAPPED_MAX || !h->h_result) return nullptr;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, h->h_result, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<double *>(dp);
}

// peer.cu: in-place sum over the ranks of n 8-byte elements (double, or unsigned long long with as_u64) at data_dev,
// asynchronous on the handle's stream; a collective (needs b200mc_peer_connect)
int peer_allreduce_async(b200mc_handle *h, void *data_dev, int32_t n, bool as_u64);
// non-zero (B200MC_ECUDA + message) once an exchange on this handle has timed out; call after a stream synchronisation
int peer_check(b200mc_handle *h);

// pcg64.cu: n outputs of NumPy's PCG64 .random() starting at output index `first` -> out_dev; st = {state_hi, state_lo,
// inc_hi, inc_lo}; `chunk` consecutive outputs per thread
int pcg64_uniform_async(b200mc_handle *h, const uint64_t st[4], uint64_t first, int64_t n, int chunk, double *out_dev);

// ---- warp / block reduction of NV doubles, deterministic "last block sums the partials" finish ---------
template <int NV>
__device__ __forceinline__ void warp_reduce(double (&v)[NV])
{
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
}

// Block partial -> partials[blockIdx.x * NV + i]; the last block to arrive adds all partials in block
// order (fixed order => bitwise reproducible for a given launch geometry) and writes out[i].
// smem must hold (blockDim.x / 32) * NV doubles.
template <int NV>
__device__ __forceinline__ void block_finish(double (&v)[NV], double *smem, double *partials,
                                             unsigned int *counter, double *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    warp_reduce<NV>(v);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[warp * NV + i] = v[i];
    }
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += smem[w * NV + threadIdx.x];
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        // every thread adds a strided subset of the CTA partials (loads issued back to back), then the subsets are
        // folded warp by warp in a fixed order -- one thread walking gridDim.x dependent L2 round trips cost ~0.3 ms
        __threadfence();
        double t[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) t[i] = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] += partials[(size_t)b * NV + i];
        }
        warp_reduce<NV>(t);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) smem[warp * NV + i] = t[i];
        }
        __syncthreads();
        if (threadIdx.x < NV) {
            double s = 0.0;
            for (int w = 0; w < nwarp; ++w) s += smem[w * NV + threadIdx.x];
            out[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;   // re-arm for the next launch on this stream
    }
}

} // namespace b200mc
