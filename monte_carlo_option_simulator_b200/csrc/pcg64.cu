// pcg64.cu -- NumPy's default_rng(seed).random(...) reproduced on the device, bit for bit.
//
// The reference draws its jump uniforms with np.random.default_rng(seed + 1).random((n_paths, n_steps))
// (engine/monte_carlo.py:308): PCG64 = pcg64 XSL-RR 128/64, state <- state * MULT + inc (mod 2^128), output =
// rotr64(hi ^ lo, state >> 122) of the NEW state, double = (out >> 11) * 2^-53.  The stream is sequential, but an LCG
// can be advanced by k steps in O(log k): thread i jumps to output index i * chunk with the standard
// "square and multiply" on (multiplier, increment) pairs in 128-bit arithmetic and then produces its chunk.  The caller
// passes the generator's 128-bit state and increment (NumPy: default_rng(seed).bit_generator.state), so the seeding hash
// (SeedSequence) stays in NumPy.  With this the reference's use_sobol=True front end needs no host array at all.
#include "common.cuh"

namespace b200mc {

struct U128 {
    unsigned long long hi, lo;
};
__host__ __device__ __forceinline__ U128 mul128(U128 a, U128 b)
{
    U128 r;
#ifdef __CUDA_ARCH__
    r.lo = a.lo * b.lo;
    r.hi = __umul64hi(a.lo, b.lo) + a.hi * b.lo + a.lo * b.hi;
#else
    const unsigned __int128 p = ((unsigned __int128)a.hi << 64 | a.lo) * ((unsigned __int128)b.hi << 64 | b.lo);
    r.lo = (unsigned long long)p;
    r.hi = (unsigned long long)(p >> 64);
#endif
    return r;
}
__host__ __device__ __forceinline__ U128 add128(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}

struct Pcg64Args {
    U128 state, inc;
    unsigned long long first;      // index of the first output wanted (0 = the generator's next output)
    long long n;                   // outputs wanted
    int chunk;                     // consecutive outputs per thread
};

__global__ void __launch_bounds__(256) k_pcg64_uniform(const __grid_constant__ Pcg64Args a, double *__restrict__ out)
{
    const U128 MULT = {0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull};
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * a.chunk;
    if (i0 >= a.n) return;
    // advance the state by (first + i0) steps
    unsigned long long delta = a.first + (unsigned long long)i0;
    U128 acc_mult = {0ull, 1ull}, acc_plus = {0ull, 0ull}, cur_mult = MULT, cur_plus = a.inc;
    while (delta) {
        if (delta & 1ull) {
            acc_mult = mul128(acc_mult, cur_mult);
            acc_plus = add128(mul128(acc_plus, cur_mult), cur_plus);
        }
        cur_plus = mul128(add128(cur_mult, U128{0ull, 1ull}), cur_plus);
        cur_mult = mul128(cur_mult, cur_mult);
        delta >>= 1;
    }
    U128 s = add128(mul128(acc_mult, a.state), acc_plus);
    const long long i1 = min(a.n, i0 + a.chunk);
    for (long long i = i0; i < i1; ++i) {
        s = add128(mul128(s, MULT), a.inc);
        const unsigned long long x = s.hi ^ s.lo;
        const unsigned int rot = (unsigned int)(s.hi >> 58);
        const unsigned long long o = (x >> rot) | (x << ((64u - rot) & 63u));
        out[i] = (double)(o >> 11) * (1.0 / 9007199254740992.0);
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
