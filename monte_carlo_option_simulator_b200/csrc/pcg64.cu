// pcg64.cu -- NumPy's default_rng(seed).random(...) reproduced on the device, bit for bit.
//
// The reference draws its jump uniforms with np.random.default_rng(seed + 1).random((n_paths, n_steps))
// (engine/monte_carlo.py:308): PCG64 = pcg64 XSL-RR 128/64, state <- state * MULT + inc (mod 2^128), output =
// rotr64(hi ^ lo, state >> 122) of the NEW state, double = (out >> 11) * 2^-53.  The stream is sequential, but an LCG
// can be advanced by k steps in O(log k): thread i jumps to output index i * chunk with the standard
// "square and multiply" on (multiplier, increment) pairs in 128-bit arithmetic and then produces its chunk.  The caller
// passes the generator's 128-bit state and increment (NumPy: default_rng(seed).bit_generator.state), so the seeding hash
// (SeedSequence) stays in NumPy.  With this the reference's use_sobol=True front end needs no host array at all.
#include "pcg64.cuh"

namespace b200mc {

struct Pcg64Args {
    U128 state, inc;
    unsigned long long first;      // index of the first output wanted (0 = the generator's next output)
    long long n;                   // outputs wanted
    int chunk;                     // consecutive outputs per thread
};

__global__ void __launch_bounds__(256) k_pcg64_uniform(const __grid_constant__ Pcg64Args a, double *__restrict__ out)
{
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * a.chunk;
    if (i0 >= a.n) return;
    U128 s = pcg64_advance(a.state, a.inc, a.first + (unsigned long long)i0);      // jump to output index first + i0
    const long long i1 = min(a.n, i0 + a.chunk);
    for (long long i = i0; i < i1; ++i) {
        out[i] = pcg64_double(pcg64_next(s, a.inc));
    }
}

// n uniforms starting at output index `first` of the generator -> out_dev (device), asynchronous.
int pcg64_uniform_async(b200mc_handle *h, const uint64_t st[4], uint64_t first, int64_t n, int chunk, double *out_dev)
{
    Pcg64Args a;
    a.state = {st[0], st[1]};
    a.inc = {st[2], st[3]};
    a.first = first;
    a.n = n;
    a.chunk = chunk < 1 ? 1 : chunk;
    const int64_t threads = (n + a.chunk - 1) / a.chunk;
    k_pcg64_uniform<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(a, out_dev);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_pcg64_random(b200mc_handle *h, const uint64_t state[4], uint64_t first, int64_t n, double *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!state || !out || n <= 0) return fail(h, B200MC_EINVAL, "state / out must be given and n positive");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n * 8));
    B200MC_TRY(pcg64_uniform_async(h, state, first, n, 64, (double *)h->d_stage));
    B200MC_CUDA(h, cudaMemcpyAsync(out, h->d_stage, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
