// philox.cuh -- counter-based Philox4x32-10 and the uniform -> normal transforms, in registers.
//
// Not in the reference: this replaces its host RNG front end (engine/monte_carlo.py:301-308,
// np.random.default_rng(...).standard_normal / .random) for the fused modes.  Layout and transforms are
// part of the ABI contract (include/b200mc.h, "Random numbers") because b200mc_dump_normals must return
// the very same values so the reference can be fed identical draws.
//
// Cost model (SASS, sm_100a, measured with csrc/microbench.cu): one round = 2 IMAD.WIDE.U32 + 2 LOP3 (round keys
// are kernel parameters, i.e. uniform-register operands of the LOP3).  IMAD.WIDE holds the sub-partition's issue
// port for 4 cycles, so the multiplies ARE the cost of a call (~68 of ~90 issue cycles): the streams below
// therefore squeeze 8 normals out of every call (one Box-Muller pair per 32-bit word).
#pragma once
#include <stdint.h>

namespace b200mc {

struct PhiloxKey {          // the 10 round keys, precomputed on the host (uniform across the grid)
    uint32_t k0[10];
    uint32_t k1[10];
};

__host__ inline PhiloxKey philox_make_key(uint64_t seed)
{
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        k.k0[i] = a;
        k.k1[i] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const PhiloxKey &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return U4{c0, c1, c2, c3};
}

// ---- approximate special functions: one MUFU each ------------------------------------------------------
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x)
{
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// float in [1,2) from the low 23 bits of a word: one LOP3, no I2F
__device__ __forceinline__ float mant12(uint32_t w)
{
    return __uint_as_float((w & 0x007fffffu) | 0x3f800000u);
}

// sqrt(2 ln 2): the Box-Muller radius is computed as sqrt(-lg2(u1)); this constant restores the scale.
// The fused kernels fold it into their per-step weight, b200mc_dump_normals multiplies it in (in double).
#define B200MC_BM_SCALE 1.1774100225154747

// Box-Muller on ONE word: a 32-bit word yields a whole pair of normals.
//   radius  u1    = 2 - mant12(w)                 in (0, 1]    low 23 bits of w  (lg2 <= 0 and never -inf)
//           rad   = sqrt(-lg2 u1)                              max radius sqrt(2*23*ln 2) = 5.65 sigma
//   angle   f2    = float with mantissa (w >> 9)  in [1, 2)    top 23 bits of w, one funnel shift
//           ang   = 2 pi f2 - 3 pi                in [-pi, pi)  (best range of sin/cos.approx)
// The two fields share bits 9..22.  That is harmless: for any fixed radius field the angle still runs over 512
// equally spaced directions (bits 23..31) plus an offset, so E[g(rad) e^{ik ang}] = 0 for every k that is not
// a multiple of 512 -- all joint moments of the pair up to order 511 are those of two independent normals,
// exactly as for a continuous angle.  Returns the UNSCALED pair (rc, rs) = rad * (cos, sin)(ang): the standard
// normals are B200MC_BM_SCALE * rc and B200MC_BM_SCALE * rs.  Cost: 2 ALU + 3 FP32 + 4 MUFU per pair.
struct BM2 { float rc, rs; };
__device__ __forceinline__ BM2 box_muller_word(uint32_t w)
{
    const float u1 = 2.0f - mant12(w);
    const float rad = sqrt_approx(-lg2_approx(u1));
    const float f2 = __uint_as_float(__funnelshift_r(w, 0x7Fu, 9));      // (w >> 9) | 0x3f800000
    const float ang = fmaf(f2, 6.283185307179586f, -9.42477796076938f);
    BM2 o;
    o.rc = rad * cos_approx(ang);
    o.rs = rad * sin_approx(ang);
    return o;
}

// Validation twin (B200MC_WIDE_RNG, never the default): the SAME transform fed by TWO words -- the radius from the low 23
// bits of one, the angle from the top 23 bits of the other -- so the two fields share no bit (4 normals per Philox call
// instead of 8).  It exists to show that the bit sharing of the production layout is invisible to the path functionals:
// tests/test_gpu_rng_quality.py prices the same options with both layouts at 1e9 paths.
__device__ __forceinline__ BM2 box_muller_wide(uint32_t w_radius, uint32_t w_angle)
{
    const float u1 = 2.0f - mant12(w_radius);
    const float rad = sqrt_approx(-lg2_approx(u1));
    const float f2 = __uint_as_float(__funnelshift_r(w_angle, 0x7Fu, 9));
    const float ang = fmaf(f2, 6.283185307179586f, -9.42477796076938f);
    BM2 o;
    o.rc = rad * cos_approx(ang);
    o.rs = rad * sin_approx(ang);
    return o;
}

// Same transform with the three factors kept apart: normals = B200MC_BM_SCALE * rad * (cs, sn).  The stochastic-variance
// kernels (fp32 state) fold their per-step constants into `rad` once instead of scaling both products per state.
struct BM3 { float rad, cs, sn; };
__device__ __forceinline__ BM3 box_muller_parts(uint32_t w)
{
    const float u1 = 2.0f - mant12(w);
    const float f2 = __uint_as_float(__funnelshift_r(w, 0x7Fu, 9));
    const float ang = fmaf(f2, 6.283185307179586f, -9.42477796076938f);
    BM3 o;
    o.rad = sqrt_approx(-lg2_approx(u1));
    o.cs = cos_approx(ang);
    o.sn = sin_approx(ang);
    return o;
}

// uniform in (0, 1) on the 2^32 grid, as a double: U = (w + 0.5) / 2^32
__device__ __forceinline__ double word_uniform(uint32_t w)
{
    return ((double)w + 0.5) * 2.3283064365386963e-10;
}

// ---- jumps (SVJ): the gap to the next jump instead of one Bernoulli test per step --------------------------------------
// The reference tests U < p = lambda_j dt at EVERY step (engine/monte_carlo.py:233): the number of jump-free steps before
// the next jump is geometric, P(G = g) = (1 - p)^g p.  Drawing G = floor(ln U / ln(1 - p)) from ONE uniform per JUMP is
// the same process in distribution and takes the jump uniform out of the per-step loop (it cost half a Philox call per
// step).  inv_lg2_q = 1 / lg2(1 - p) (<= 0; 0 when p >= 1: a jump at every step), computed on the host in double.
// Everything here is fp32 for every path-state precision, so the fp32 / fp64 kernels and b200mc_dump_normals see the
// same jump times.
__device__ __forceinline__ int jump_gap(uint32_t w, float inv_lg2_q)
{
    const float u = ((float)w + 0.5f) * 2.3283064365386963e-10f;           // (0, 1]
    const float g = floorf(log2f(u) * inv_lg2_q);                          // >= 0 (or -0)
    return (int)fminf(g, 1.0e9f);
}

struct JumpStream {
    // Philox stream B200MC_STREAM_SVJ_JUMP, counter block k -> jumps 2k and 2k + 1 of the path:
    //   (w0 -> gap before jump 2k, w1 -> its size), (w2 -> gap before jump 2k + 1, w3 -> its size)
    // size: Z_jump_size = B200MC_BM_SCALE * rc(BM(w)) (the cosine member of the word's Box-Muller pair)
    // (Tried: the next 2-4 jumps of a path read ahead into a small ring before the step loop, so that the divergent branch
    // never holds a Philox call -- 4.30e11 -> 4.02e11 path-steps/s on B200: the read-ahead and the ring's selects cost
    // more than the rare refills they remove.)
    int next, next2;            // step index of the next jump and of the one after it
    uint32_t w_size, w_size2;   // their size words
    uint32_t m;                 // jumps consumed so far

    __device__ __forceinline__ void refill(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, int from)
    {
        const U4 q = philox4x32_10(c0, c1, m >> 1, B200MC_STREAM_SVJ_JUMP, key);
        next = from + jump_gap(q.x, inv_lg2_q);
        w_size = q.y;
        next2 = next + 1 + jump_gap(q.z, inv_lg2_q);
        w_size2 = q.w;
    }
    __device__ __forceinline__ void init(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, bool on)
    {
        m = 0u;
        next = next2 = 0x7fffffff;
        w_size = w_size2 = 0u;
        if (on) refill(c0, c1, key, inv_lg2_q, 0);
    }
    // the jump at step `s` (== next) has been applied: move to the following one
    __device__ __forceinline__ void advance(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, int s)
    {
        ++m;
        if (m & 1u) { next = next2; w_size = w_size2; }
        else refill(c0, c1, key, inv_lg2_q, s + 1);
    }
    // unscaled size draw of the pending jump: Z_jump_size = B200MC_BM_SCALE * size_raw()
    __device__ __forceinline__ float size_raw() const { return box_muller_word(w_size).rc; }
};

} // namespace b200mc
