// dump.cu -- the draws of the fused kernels, exported so that the reference (or the oracle) can be fed the
// IDENTICAL numbers ("deterministic mode" of the acceptance contract).  Not in the reference.
//
// b200mc_dump_normals returns, as float64 [n_paths, n_steps], exactly the values the fused kernels consume:
//   Z1 / Z2        BM_SCALE * (double)raw   with raw the fp32 Box-Muller output of philox.cuh
//   Z_jump         (SVJ) the fused kernels draw jump TIMES (geometric gaps, philox.cuh JumpStream), the reference wants
//                  a uniform per step that it compares with jump_prob = lambda_j dt (monte_carlo.py:233).  The array
//                  returned here is i.i.d. U(0,1) in distribution and fires exactly at the kernels' jump steps:
//                  jump_prob * U' at a jump step (U' from the jump's size word), jump_prob + (1 - jump_prob) * U''
//                  elsewhere (U'' from B200MC_STREAM_FILL, block s / 4, word s % 4).  1.0 = "never jumps" on the
//                  other streams.
//   Z_jump_size    BM_SCALE * (double)raw of the jump's size word at the jump steps, 0 elsewhere -- the reference only
//                  reads it where the jump fires (monte_carlo.py:233-234)
// Arrays a stream does not carry come back as neutral values (Z2 = 0, Z_jump = 1, Z_jump_size = 0).
#include "prep.cuh"

namespace b200mc {

__global__ void k_dump_philox(PhiloxKey key, uint64_t path0, int64_t n_paths, int n_blocks, uint32_t stream,
                              uint32_t *__restrict__ out)
{
    const int64_t total = n_paths * n_blocks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t path = path0 + (uint64_t)(i / n_blocks);
        const uint32_t blk = (uint32_t)(i % n_blocks);
        const U4 w = philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), blk, stream, key);
        reinterpret_cast<uint4 *>(out)[i] = make_uint4(w.x, w.y, w.z, w.w);
    }
}

__global__ void k_dump_normals(PhiloxKey key, uint64_t path0, int64_t n_paths, int n_steps, uint32_t stream, int which,
                               double *__restrict__ out)
{
    const bool gbm_layout = stream == B200MC_STREAM_GBM || stream == B200MC_STREAM_HEDGE;
    const int per = gbm_layout ? 8 : 4;
    const int n_blocks = (n_steps + per - 1) / per;
    const int64_t total = n_paths * n_blocks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pi = i / n_blocks;
        const uint64_t path = path0 + (uint64_t)pi;
        const uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32);
        const int blk = (int)(i % n_blocks);
        const U4 w = philox4x32_10(c0, c1, (uint32_t)blk, stream, key);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
        const double neutral = which == B200MC_ZJUMP_U ? 1.0 : 0.0;
        double vals[8];
        if (gbm_layout) {
            for (int t = 0; t < 4; ++t) {
                const BM2 b = box_muller_word(ww[t]);
                vals[2 * t] = which == B200MC_Z1 ? B200MC_BM_SCALE * (double)b.rc : neutral;
                vals[2 * t + 1] = which == B200MC_Z1 ? B200MC_BM_SCALE * (double)b.rs : neutral;
            }
        } else {                       // HESTON and the diffusion of SVJ (the jump arrays come from k_dump_jumps)
            for (int t = 0; t < 4; ++t) {
                const BM2 b = box_muller_word(ww[t]);
                vals[t] = which == B200MC_Z1 ? B200MC_BM_SCALE * (double)b.rc
                        : which == B200MC_Z2 ? B200MC_BM_SCALE * (double)b.rs : neutral;
            }
        }
        for (int t = 0; t < per; ++t) {
            const int s = blk * per + t;
            if (s < n_steps) out[(size_t)pi * n_steps + s] = vals[t];
        }
    }
}

// The jump arrays of the SVJ model: one thread walks one path's jump stream exactly as simulate_path does.
__global__ void k_dump_jumps(PhiloxKey key, uint64_t path0, int64_t n_paths, int n_steps, int which, double jump_prob,
                             float inv_lg2_q, int jump_on, double *__restrict__ out)
{
    for (int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pi < n_paths; pi += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t path = path0 + (uint64_t)pi;
        const uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32);
        JumpStream jmp;
        jmp.init(c0, c1, key, inv_lg2_q, jump_on != 0);
        U4 fill = U4{0u, 0u, 0u, 0u};
        for (int s = 0; s < n_steps; ++s) {
            if (which == B200MC_ZJUMP_U && (s & 3) == 0)
                fill = philox4x32_10(c0, c1, (uint32_t)(s >> 2), B200MC_STREAM_FILL, key);
            const uint32_t fw = (s & 3) == 0 ? fill.x : (s & 3) == 1 ? fill.y : (s & 3) == 2 ? fill.z : fill.w;
            double val;
            if (s == jmp.next) {
                if (which == B200MC_ZJUMP_U) val = jump_prob * word_uniform(jmp.w_size);      // < jump_prob
                else val = B200MC_BM_SCALE * (double)jmp.size_raw();
                jmp.advance(c0, c1, key, inv_lg2_q, s);
            } else if (which == B200MC_ZJUMP_U) {
                val = jump_prob + (1.0 - jump_prob) * word_uniform(fw);
                if (!(val > jump_prob)) val = 1.0;                                            // never below the threshold
            } else {
                val = 0.0;
            }
            out[(size_t)pi * n_steps + s] = val;
        }
    }
}

// Raw moments of the normals of the GBM stream over paths x blocks (8 draws per block), fp64 accumulation: the
// statistical certificate of the generator (tools/normal_moments.py).  out[0..5] = sum z^k, k = 0..4, and sum z_a z_b
// over the two members of every Box-Muller pair.
__global__ void __launch_bounds__(256)
k_normal_moments(const __grid_constant__ PhiloxKey key, uint64_t path0, int64_t n_paths, int n_blocks, double *partials,
                 unsigned int *counter, double *out)
{
    __shared__ double smem[(256 / 32) * 6];
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * 256) {
        const uint64_t path = path0 + (uint64_t)i;
        for (int j = 0; j < n_blocks; ++j) {
            const U4 w = philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)j, B200MC_STREAM_GBM, key);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const BM2 b = box_muller_word(ww[t]);
                const double a = B200MC_BM_SCALE * (double)b.rc, c = B200MC_BM_SCALE * (double)b.rs;
                const double a2 = a * a, c2 = c * c;
                v[0] += 2.0;
                v[1] += a + c;
                v[2] += a2 + c2;
                v[3] += a2 * a + c2 * c;
                v[4] += a2 * a2 + c2 * c2;
                v[5] += a * c;
            }
        }
    }
    block_finish<6>(v, smem, partials, counter, out);
}

// Joint distribution of consecutive normals of the GBM stream on a 64 x 64 grid of EQUIPROBABLE cells (cell = floor(64
// Phi(z))): the input of the chi-square / Kolmogorov checks of the generator (tests/test_gpu_rng_quality.py).
//   lag 0: the two members of every Box-Muller pair (same 32-bit word)
//   lag 1: the second member of word i with the first member of word i + 1 (adjacent words of a Philox block)
__global__ void __launch_bounds__(256)
k_normal_hist2d(const __grid_constant__ PhiloxKey key, uint64_t path0, int64_t n_paths, int n_blocks, int lag,
                unsigned long long *__restrict__ out)
{
    __shared__ unsigned int hist[64 * 64];
    for (int i = threadIdx.x; i < 4096; i += 256) hist[i] = 0u;
    __syncthreads();
    auto cell = [](float raw) -> int {
        const double u = normcdf(B200MC_BM_SCALE * (double)raw);
        const int c = (int)(u * 64.0);
        return c > 63 ? 63 : c;
    };
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * 256) {
        const uint64_t path = path0 + (uint64_t)i;
        for (int j = 0; j < n_blocks; ++j) {
            const U4 w = philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)j, B200MC_STREAM_GBM, key);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            BM2 b[4];
            if (lag & B200MC_HIST_WIDE) {
                b[0] = box_muller_wide(ww[0], ww[1]);
                b[1] = box_muller_wide(ww[2], ww[3]);
                if ((lag & 1) == 0) {
                    atomicAdd(&hist[cell(b[0].rc) * 64 + cell(b[0].rs)], 1u);
                    atomicAdd(&hist[cell(b[1].rc) * 64 + cell(b[1].rs)], 1u);
                } else {
                    atomicAdd(&hist[cell(b[0].rs) * 64 + cell(b[1].rc)], 1u);
                }
                continue;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = (lag & B200MC_HIST_R01) ? box_muller_word_r01(ww[t]) : box_muller_word(ww[t]);
            if ((lag & 1) == 0) {
#pragma unroll
                for (int t = 0; t < 4; ++t) atomicAdd(&hist[cell(b[t].rc) * 64 + cell(b[t].rs)], 1u);
            } else {
#pragma unroll
                for (int t = 0; t < 3; ++t) atomicAdd(&hist[cell(b[t].rs) * 64 + cell(b[t + 1].rc)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += 256)
        if (hist[i]) atomicAdd(&out[i], (unsigned long long)hist[i]);
}

} // namespace b200mc
using namespace b200mc;

extern "C" int b200mc_normal_hist2d(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths, int32_t n_blocks,
                                    int lag, uint64_t out[4096])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || n_paths <= 0 || n_blocks <= 0 || (lag & ~(1 | B200MC_HIST_WIDE | B200MC_HIST_R01))) return fail(h, B200MC_EINVAL, "bad argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    // a CTA's shared 32-bit cells must not overflow: at most 2^31 pairs per CTA
    const int grid = h->sm_count * 8;
    if ((double)n_paths * n_blocks * 4.0 / grid > 2.0e9) return fail(h, B200MC_EINVAL, "too many draws per launch: split the call");
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, 4096 * 8));
    B200MC_CUDA(h, cudaMemsetAsync(h->d_scratch, 0, 4096 * 8, h->stream));
    k_normal_hist2d<<<grid, 256, 0, h->stream>>>(philox_make_key(seed), path_offset, n_paths, n_blocks, lag,
                                                 (unsigned long long *)h->d_scratch);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, 4096 * 8, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mc_normal_moments(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths,
                                     int32_t n_blocks, double out[6])
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || n_paths <= 0 || n_blocks <= 0) return fail(h, B200MC_EINVAL, "bad argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const int grid = h->sm_count * 8;
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, (size_t)(grid + 1) * 6 * 8));
    double *partials = (double *)h->d_scratch, *res = partials + (size_t)grid * 6;
    k_normal_moments<<<grid, 256, 0, h->stream>>>(philox_make_key(seed), path_offset, n_paths, n_blocks, partials,
                                                  h->d_counter, res);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, res, 48, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mc_dump_philox(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths,
                                  int32_t n_blocks, uint32_t stream, uint32_t *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || n_paths <= 0 || n_blocks <= 0) return fail(h, B200MC_EINVAL, "bad argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)n_paths * n_blocks * 16;
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, bytes));
    k_dump_philox<<<h->sm_count * 4, 256, 0, h->stream>>>(philox_make_key(seed), path_offset, n_paths, n_blocks, stream,
                                                          (uint32_t *)h->d_stage);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mc_dump_normals(b200mc_handle *h, uint64_t seed, uint64_t path_offset, int64_t n_paths,
                                   int32_t n_steps, uint32_t stream, int which, double jump_prob, double *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!out || n_paths <= 0 || n_steps <= 0) return fail(h, B200MC_EINVAL, "bad argument");
    if (stream > B200MC_STREAM_HEDGE) return fail(h, B200MC_EINVAL, "unknown stream");
    if (which < B200MC_Z1 || which > B200MC_ZJUMP_SIZE) return fail(h, B200MC_EINVAL, "unknown array selector");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)n_paths * n_steps * 8;
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, bytes));
    if (stream == B200MC_STREAM_SVJ && (which == B200MC_ZJUMP_U || which == B200MC_ZJUMP_SIZE)) {
        double inv = 0.0;
        const int on = jump_setup(jump_prob, &inv);
        k_dump_jumps<<<h->sm_count * 4, 128, 0, h->stream>>>(philox_make_key(seed), path_offset, n_paths, n_steps, which,
                                                            jump_prob, (float)inv, on, (double *)h->d_stage);
    } else {
        k_dump_normals<<<h->sm_count * 4, 256, 0, h->stream>>>(philox_make_key(seed), path_offset, n_paths, n_steps, stream,
                                                               which, (double *)h->d_stage);
    }
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    B200MC_CUDA(h, cudaMemcpyAsync(out, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}


// Which draw stream a fused call with these parameters consumes (the mode selection of prep.cuh), so that a caller of
// b200mc_dump_normals can ask instead of re-deriving the rule.
extern "C" int b200mc_select_stream(b200mc_handle *h, const b200mc_svj_params *p, double T, int32_t n_steps, uint32_t flags,
                                    const b200mc_bumps *bumps, uint32_t *stream)
{
    if (!stream) return fail(h, B200MC_EINVAL, "stream is NULL");
    Prep pr;
    B200MC_TRY(prepare(h, p, 1.0, T, n_steps, 1, 0, flags, bumps, pr));
    *stream = pr.mode == MODE_SVJ ? B200MC_STREAM_SVJ : (pr.mode == MODE_HESTON ? B200MC_STREAM_HESTON : B200MC_STREAM_GBM);
    return 0;
}
