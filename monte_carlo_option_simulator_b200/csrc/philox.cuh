// philox.cuh -- counter-based Philox4x32-10 and the uniform -> normal transforms, in registers.
//
// Not in the reference: this replaces its host RNG front end (engine/monte_carlo.py:301-308,
// np.random.default_rng(...).standard_normal / .random) for the fused modes.  Layout and transforms are
// part of the ABI contract (include/b200mc.h, "Random numbers") because b200mc_dump_normals must return
// the very same values so the reference can be fed identical draws.
//
// Cost model (SASS, sm_100a): one round = 2 IMAD.WIDE.U32 + 2 LOP3 (round keys are kernel parameters,
// i.e. constant-bank operands of the LOP3), so one call = 40 integer ops for 4 words = 10 per draw.
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

// Box-Muller on two words.  Returns the UNSCALED pair (rc, rs) = sqrt(-lg2 u1) * (cos, sin)(angle):
// the standard normals are B200MC_BM_SCALE * rc and B200MC_BM_SCALE * rs.
//   u1    = 2 - mant12(wa)            in (0, 1]           (so lg2 <= 0 and never -inf)
//   angle = 2 pi (mant12(wb) - 1.5)   in [-pi, pi)        (best range of sin/cos.approx)
// 4 MUFU (lg2, sqrt, sin, cos) per pair.
struct BM2 { float rc, rs; };
__device__ __forceinline__ BM2 box_muller_raw(uint32_t wa, uint32_t wb)
{
    const float u1 = 2.0f - mant12(wa);
    const float rad = sqrt_approx(-lg2_approx(u1));
    const float ang = fmaf(mant12(wb), 6.283185307179586f, -9.42477796076938f);
    BM2 o;
    o.rc = rad * cos_approx(ang);
    o.rs = rad * sin_approx(ang);
    return o;
}

// jump uniform as the reference sees it (float64 in (0,1)): U = (w + 0.5) / 2^32
__device__ __forceinline__ double jump_uniform(uint32_t w)
{
    return ((double)w + 0.5) * 2.3283064365386963e-10;
}

// jump-size standard normal from one word (only evaluated when a jump fires): inverse normal CDF of a
// 24-bit uniform strictly inside (0,1)
__device__ __forceinline__ float jump_size_normal(uint32_t w)
{
    const float u = ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f;
    return normcdfinvf(u);
}

} // namespace b200mc
