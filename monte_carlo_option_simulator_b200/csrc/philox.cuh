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
    // The next `pf` (2 or 4) jumps of the path are kept in a small ring (thread-local memory, touched only when a jump
    // fires): the Philox calls and log2 evaluations that fill it run for ALL lanes at the start of the path, so the
    // divergent branch taken when a lane jumps is a dozen instructions instead of a Philox call -- with lambda T ~ 1
    // a path rarely needs a refill inside the step loop.  pf only sets how far ahead the stream is read, never WHAT it
    // holds (the host picks it from lambda T).
    int next;                   // step index of the next jump
    uint32_t m;                 // jumps consumed so far
    int jstep[4];               // ring: step index of jump m + i (slot (m + i) & (pf - 1))
    uint32_t jword[4];          // ring: its size word

    __device__ __forceinline__ void refill(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, int pf, int from)
    {
        int at = from;
        for (int c = 0; c < (pf >> 1); ++c) {
            const U4 q = philox4x32_10(c0, c1, (m >> 1) + (uint32_t)c, B200MC_STREAM_SVJ_JUMP, key);
            at += jump_gap(q.x, inv_lg2_q);
            jstep[2 * c] = at;
            jword[2 * c] = q.y;
            at += 1 + jump_gap(q.z, inv_lg2_q);
            jstep[2 * c + 1] = at;
            jword[2 * c + 1] = q.w;
            at += 1;
        }
        next = jstep[0];
    }
    __device__ __forceinline__ void init(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, int pf, bool on)
    {
        m = 0u;
        next = 0x7fffffff;
        if (on) refill(c0, c1, key, inv_lg2_q, pf, 0);
    }
    // unscaled size draw of the pending jump: Z_jump_size = B200MC_BM_SCALE * size_raw()
    __device__ __forceinline__ float size_raw(int pf) const { return box_muller_word(jword[m & (uint32_t)(pf - 1)]).rc; }
    // the jump at step `s` (== next) has been applied: move to the following one
    __device__ __forceinline__ void advance(uint32_t c0, uint32_t c1, const PhiloxKey &key, float inv_lg2_q, int pf, int s)
    {
        ++m;
        const uint32_t slot = m & (uint32_t)(pf - 1);
        if (slot == 0u) refill(c0, c1, key, inv_lg2_q, pf, s + 1);
        else next = jstep[slot];
    }
};

} // namespace b200mc
