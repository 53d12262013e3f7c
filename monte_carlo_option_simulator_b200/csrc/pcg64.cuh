// pcg64.cuh -- NumPy's PCG64 bit generator (pcg64 XSL-RR 128/64) in 128-bit device arithmetic: state <- state * MULT + inc
// (mod 2^128), output = rotr64(hi ^ lo, state >> 122) of the NEW state; O(log k) jump-ahead.  Shared by pcg64.cu (uniforms)
// and np_normal.cu (Ziggurat normals).
#pragma once
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

__host__ __device__ __forceinline__ U128 pcg64_mult() { return U128{0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull}; }

// state after `delta` steps of the generator (state, inc): "square and multiply" on (multiplier, increment) pairs
__host__ __device__ __forceinline__ U128 pcg64_advance(U128 state, U128 inc, unsigned long long delta)
{
    U128 acc_mult = {0ull, 1ull}, acc_plus = {0ull, 0ull}, cur_mult = pcg64_mult(), cur_plus = inc;
    while (delta) {
        if (delta & 1ull) {
            acc_mult = mul128(acc_mult, cur_mult);
            acc_plus = add128(mul128(acc_plus, cur_mult), cur_plus);
        }
        cur_plus = mul128(add128(cur_mult, U128{0ull, 1ull}), cur_plus);
        cur_mult = mul128(cur_mult, cur_mult);
        delta >>= 1;
    }
    return add128(mul128(acc_mult, state), acc_plus);
}

// one step: advances s and returns the 64-bit output (NumPy's next_uint64)
__host__ __device__ __forceinline__ unsigned long long pcg64_next(U128 &s, U128 inc)
{
    s = add128(mul128(s, pcg64_mult()), inc);
    const unsigned long long x = s.hi ^ s.lo;
    const unsigned int rot = (unsigned int)(s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
// NumPy's next_double: 53 random bits
__host__ __device__ __forceinline__ double pcg64_double(unsigned long long o) { return (double)(o >> 11) * (1.0 / 9007199254740992.0); }

} // namespace b200mc
