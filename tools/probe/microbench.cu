// microbench.cu -- issue-rate probes for the pipes the fused kernel lives on.  MEASUREMENT TOOL, not product: built into its
// own library (tools/probe/libb200mc_probe.so, tools/probe/Makefile) with its own two-function C ABI; libb200mc.so neither
// contains nor links it.  It includes the product's philox.cuh only to time the very Philox / Box-Muller code that ships.  The roofline of a kernel that moves
// no data is an instruction-throughput roofline; its denominators (FP32 FMA, 32x32->64 integer multiply, 3-input
// logic, MUFU) are not in MEASURED_PEAKS.json, so bench.py measures them on the same device, in the same run,
// with these kernels ("the builder must microbenchmark FFMA, IMAD.WIDE, LOP3, MUFU").
// Each thread runs 8 independent dependency chains of one instruction type; with 2048 threads per SM resident the
// pipe, not latency, is the limit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mc.h"
#include "../../monte_carlo_option_simulator_b200/csrc/philox.cuh"

namespace b200mc {

enum { MB_FFMA = 0, MB_IMAD_WIDE = 1, MB_LOP3 = 2, MB_MUFU_EX2 = 3, MB_MUFU_SIN = 4, MB_IADD3 = 5, MB_PHILOX = 6,
       MB_PHILOX_BM = 7, MB_FMUL = 8, MB_MUFU_LG2 = 9, MB_MUFU_SQRT = 10, MB_MIX_FFMA_LOP3 = 11, MB_IMAD_LO = 12, MB_IMAD_HI = 13, MB_IMAD_LOHI = 14, MB_FFMA2 = 15, MB_F2F = 16, MB_DADD = 17,
       MB_MIX_FFMA2_LOP3 = 18, MB_MIX_FFMA2_MUFU = 19, MB_MIX_FFMA2_WIDE = 20, MB_MIX_FFMA2_FFMA = 21, MB_COUNT = 22 };

template <int WHICH>
__global__ void __launch_bounds__(256) k_microbench(int iters, uint32_t seed, const __grid_constant__ PhiloxKey key,
                                                    uint32_t *sink)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (WHICH == MB_FFMA || WHICH == MB_FMUL) {
        float a[8];
        const float b = __uint_as_float(0x3f800001u + (seed & 1)), c = __uint_as_float(0x33800000u);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (float)(t + k);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (WHICH == MB_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c));
                else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k];
        if (s == 123.456f) sink[0] = t;
    } else if constexpr (WHICH == MB_FFMA2) {
        // packed fp32 (fma.rn.f32x2 -> FFMA2): counted as INSTRUCTIONS (each does two FMAs per lane)
        unsigned long long a[8];
        const unsigned long long b = 0x3f8000013f800001ull + (seed & 1), c = 0x3380000033800000ull;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = ((unsigned long long)__float_as_uint((float)(t + k)) << 32) | __float_as_uint((float)(t + k + 1));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(b), "l"(c));
        }
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s ^= a[k];
        if (s == 0x12345ull) sink[0] = t;
    } else if constexpr (WHICH >= MB_MIX_FFMA2_LOP3 && WHICH <= MB_MIX_FFMA2_FFMA) {
        // does a packed FFMA2 hold the ISSUE port for its two FMA-pipe cycles, or only the pipe?  Four independent chains
        // of (FFMA2, partner) with the partner on another pipe: LOP3 (ALU), MUFU.EX2 (XU), IMAD.WIDE (heavy) or a plain
        // FFMA (same pipe, the control).  Counted as PAIRS.
        unsigned long long a[4];
        uint32_t u[4];
        float f[4];
        unsigned long long wd[4];
        const unsigned long long b = 0x3f8000013f800001ull + (seed & 1), c = 0x3380000033800000ull;
        const float fb = __uint_as_float(0x3f800001u + (seed & 1)), fc = __uint_as_float(0x33800000u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = ((unsigned long long)__float_as_uint((float)(t + k)) << 32) | __float_as_uint((float)(t + k + 1));
            u[k] = t * 4u + k; f[k] = 1.0f + (float)((t + k) & 7) * 0.125f; wd[k] = t + k;
        }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(b), "l"(c));
                if constexpr (WHICH == MB_MIX_FFMA2_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(seed + i), "r"(~seed));
                else if constexpr (WHICH == MB_MIX_FFMA2_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
                else if constexpr (WHICH == MB_MIX_FFMA2_WIDE) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(wd[k]) : "r"((uint32_t)(wd[k] >> 32) + (uint32_t)wd[k]), "r"(0xD2511F53u));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fb), "f"(fc));
            }
        }
        unsigned long long s = 0;
        float fs = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s ^= a[k] ^ u[k] ^ wd[k]; fs += f[k]; }
        if (s == 0x12345ull && fs == 1.5f) sink[0] = t;
    } else if constexpr (WHICH == MB_F2F) {
        // fp32 <-> fp64 conversions (the fp64 path state consumes fp32 draws): one widening + one narrowing per
        // link of the chain, counted as TWO conversions
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 1.0f + (float)((t + k) & 1023);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double d;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a[k]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a[k]) : "d"(d));
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k];
        if (s == 123.456f) sink[0] = t;
    } else if constexpr (WHICH == MB_DADD) {
        double a[8];
        const double b = 1.0 + 1e-9 * (double)(seed & 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (double)(t + k);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(b));
        }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k];
        if (s == 123.456) sink[0] = t;
    } else if constexpr (WHICH == MB_IMAD_WIDE) {
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = t * 8u + k + seed;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(a[k]));
                a[k] = (uint32_t)(p >> 32);       // hi word feeds the next multiply (a register move at most)
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s ^= a[k];
        if (s == 0x12345u) sink[0] = t;
    } else if constexpr (WHICH == MB_IMAD_LO || WHICH == MB_IMAD_HI || WHICH == MB_IMAD_LOHI) {
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = t * 8u + k + seed;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (WHICH == MB_IMAD_LO) asm volatile("mul.lo.u32 %0, %0, 0xD2511F53;" : "+r"(a[k]));
                else if constexpr (WHICH == MB_IMAD_HI) asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(a[k]));
                else {   // the pair that replaces one mul.wide: lo and hi of the same product
                    uint32_t lo, hi;
                    asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(a[k]));
                    asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(a[k]));
                    a[k] = lo ^ hi;
                }
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s ^= a[k];
        if (s == 0x12345u) sink[0] = t;
    } else if constexpr (WHICH == MB_LOP3 || WHICH == MB_IADD3) {
        uint32_t a[8];
        const uint32_t b = seed * 2654435761u, c = ~seed;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = t * 8u + k;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (WHICH == MB_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b + i), "r"(c));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s ^= a[k];
        if (s == 0x12345u) sink[0] = t;
    } else if constexpr (WHICH == MB_MUFU_EX2 || WHICH == MB_MUFU_SIN || WHICH == MB_MUFU_LG2 || WHICH == MB_MUFU_SQRT) {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 0.5f + 1e-3f * (float)((t + k) & 255);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if constexpr (WHICH == MB_MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
                else if constexpr (WHICH == MB_MUFU_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
                else if constexpr (WHICH == MB_MUFU_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
                else asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
            }
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k];
        if (s == 123.456f) sink[0] = t;
    } else if constexpr (WHICH == MB_MIX_FFMA_LOP3) {
        float a[4];
        uint32_t u[4];
        const float b = __uint_as_float(0x3f800001u + (seed & 1)), c = __uint_as_float(0x33800000u);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = (float)(t + k); u[k] = t * 4u + k; }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(seed + i), "r"(~seed));
            }
        }
        float s = 0.f;
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s += a[k]; x ^= u[k]; }
        if (s == 123.456f && x == 77u) sink[0] = t;
    } else {   // MB_PHILOX, MB_PHILOX_BM: iters Philox4x32-10 calls (+ two Box-Muller pairs) per thread
        float s = 0.f;
        uint32_t x = 0;
        for (int i = 0; i < iters; ++i) {
            const U4 w = philox4x32_10(t, seed, (uint32_t)i, 0u, key);
            if constexpr (WHICH == MB_PHILOX) x ^= w.x ^ w.y ^ w.z ^ w.w;
            else {
                const BM2 b0 = box_muller_word(w.x), b1 = box_muller_word(w.y), b2 = box_muller_word(w.z),
                          b3 = box_muller_word(w.w);
                s += b0.rc; s += b0.rs; s += b1.rc; s += b1.rs; s += b2.rc; s += b2.rs; s += b3.rc; s += b3.rs;
            }
        }
        if (s == 123.456f || x == 0x12345u) sink[0] = t;
    }
}

// Mixed probe: per iteration NW IMAD.WIDE + NL LOP3 + NM MUFU + NF FFMA over independent chains -- shows which pipes
// overlap and which share a dispatch port (the fused kernel's hot loop is such a mix).
// PACK: the NF FMAs are issued as NF/2 packed FFMA2 (fma.rn.f32x2) -- the same arithmetic in half the instructions.
template <int NW, int NL, int NM, int NF, bool PACK = false>
__global__ void __launch_bounds__(256) k_mix(int iters, uint32_t seed, uint32_t *sink)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w[8], l[8];
    float m[8], f[8];
    unsigned long long pf[8];
    const float b = __uint_as_float(0x3f800001u + (seed & 1)), c = __uint_as_float(0x33800000u);
    const unsigned long long pb = 0x3f8000013f800001ull + (seed & 1), pc = 0x3380000033800000ull;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        w[k] = t * 8u + k + seed; l[k] = t + k; m[k] = 0.5f + 1e-3f * (float)((t + k) & 255); f[k] = (float)(t + k);
        pf[k] = ((unsigned long long)__float_as_uint((float)(t + k)) << 32) | __float_as_uint((float)(t + k + 1));
    }
    for (int i = 0; i < iters; ++i) {
        constexpr int NMAX = NW > NL ? (NW > NM ? (NW > NF ? NW : NF) : (NM > NF ? NM : NF)) : (NL > NM ? (NL > NF ? NL : NF) : (NM > NF ? NM : NF));
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {
            if (k < NW) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(w[k & 7])); w[k & 7] = (uint32_t)(p >> 32); }
            if (k < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[k & 7]) : "r"(seed + i), "r"(~seed));
            if (k < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k & 7]));
            if constexpr (PACK) { if (k < NF / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pf[k & 7]) : "l"(pb), "l"(pc)); }
            else if (k < NF) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k & 7]) : "f"(b), "f"(c));
        }
    }
    float s = 0.f;
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += m[k] + f[k]; x ^= w[k] ^ l[k] ^ (uint32_t)pf[k] ^ (uint32_t)(pf[k] >> 32); }
    if (s == 123.456f && x == 77u) sink[0] = t;
}

template <int W> static void mb_launch(int grid, int iters, uint32_t seed, const PhiloxKey &key, uint32_t *sink, cudaStream_t st)
{
    k_microbench<W><<<grid, 256, 0, st>>>(iters, seed, key, sink);
}

} // namespace b200mc
using namespace b200mc;

// Mixed probe: combo selects (NW, NL, NM, NF) from a fixed table; *iters_per_s = thread-iterations per second.
namespace {
struct Probe {                      // one device context per call: stream, two events, a sink word
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t *sink = nullptr;
    int open(int device)
    {
        cudaDeviceProp prop;
        if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
        if (prop.major != 10) return 2;                                   // sm_100a code only
        sm_count = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return 1;
        if (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess) return 1;
        if (cudaMalloc(&sink, 64) != cudaSuccess) return 1;
        return 0;
    }
    ~Probe()
    {
        if (sink) cudaFree(sink);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};
#define PROBE_CUDA(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); return 3; } } while (0)
} // namespace

// Returns 0, or 1 (CUDA set-up failed), 2 (not an sm_100 device), 3 (CUDA error), 4 (bad argument).
extern "C" int b200mc_probe_mix(int device, int combo, int iters, double *iters_per_s, int counts[4])
{
    if (!iters_per_s || !counts || iters <= 0) return 4;
    Probe pb;
    if (int rc = pb.open(device)) return rc;
    Probe *h = &pb;
    const int grid = h->sm_count * 8;
    uint32_t *sink = h->sink;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PROBE_CUDA(cudaEventRecord(h->ev0, h->stream));
#define MIXCASE(id, a, b, c, d) case id: k_mix<a, b, c, d><<<grid, 256, 0, h->stream>>>(iters, rep, sink); counts[0] = a; counts[1] = b; counts[2] = c; counts[3] = d; break;
#define MIXPACK(id, a, b, c, d) case id: k_mix<a, b, c, d, true><<<grid, 256, 0, h->stream>>>(iters, rep, sink); counts[0] = a; counts[1] = b; counts[2] = c; counts[3] = d; break;
        switch (combo) {
        MIXCASE(0, 8, 0, 0, 0)
        MIXCASE(1, 0, 8, 0, 0)
        MIXCASE(2, 0, 0, 8, 0)
        MIXCASE(3, 0, 0, 0, 8)
        MIXCASE(4, 8, 0, 4, 0)
        MIXCASE(5, 8, 12, 0, 0)
        MIXCASE(6, 0, 12, 4, 0)
        MIXCASE(7, 0, 0, 4, 8)
        MIXCASE(8, 8, 0, 0, 8)
        MIXCASE(9, 8, 12, 4, 4)
        MIXCASE(10, 8, 12, 4, 0)
        MIXCASE(11, 8, 0, 4, 4)
        MIXCASE(12, 8, 12, 2, 4)
        MIXCASE(13, 6, 12, 4, 4)
        MIXCASE(14, 8, 16, 4, 8)
        MIXCASE(15, 0, 16, 0, 16)
        MIXCASE(16, 4, 8, 4, 4)
        // the fused kernels' mixes with plain and with packed FMAs (counts[3] stays the number of FMAs)
        MIXCASE(17, 8, 12, 4, 8)
        MIXPACK(18, 8, 12, 4, 8)
        MIXCASE(19, 4, 9, 6, 16)
        MIXPACK(20, 4, 9, 6, 16)
        MIXCASE(21, 2, 4, 2, 10)
        MIXPACK(22, 2, 4, 2, 10)
        default: return 4;
        }
#undef MIXCASE
#undef MIXPACK
        PROBE_CUDA(cudaGetLastError());
        PROBE_CUDA(cudaEventRecord(h->ev1, h->stream));
        PROBE_CUDA(cudaEventSynchronize(h->ev1));
        float ms = 0.f;
        PROBE_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    *iters_per_s = (double)grid * 256.0 * (double)iters / ((double)best * 1e-3);
    return 0;
}

// which: 0 FFMA, 1 IMAD.WIDE.U32, 2 LOP3, 3 MUFU.EX2, 4 MUFU.SIN, 5 IADD, 6 Philox4x32-10 calls, 7 Philox + 4 Box-Muller
// pairs = 8 normals (counted as calls), 8 FMUL, 9 MUFU.LG2, 10 MUFU.SQRT, 11 FFMA+LOP3 interleaved (counted as pairs),
// 12-14 IMAD lo / hi / lo+hi, 15 FFMA2, 16 fp32<->fp64 conversions, 17 DADD, 18-21 FFMA2 paired with LOP3 / MUFU.EX2 /
// IMAD.WIDE / FFMA (counted as pairs).
// *ops_per_s = thread-level operations per second over the whole device (kernel time by CUDA events, best of 3).
extern "C" int b200mc_probe_rate(int device, int which, int iters, double *ops_per_s)
{
    if (!ops_per_s || which < 0 || which >= MB_COUNT || iters <= 0) return 4;
    Probe pb;
    if (int rc = pb.open(device)) return rc;
    Probe *h = &pb;
    const int grid = h->sm_count * 8;
    const PhiloxKey key = philox_make_key(0x9E3779B97F4A7C15ull);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PROBE_CUDA(cudaEventRecord(h->ev0, h->stream));
        switch (which) {
        case 0: mb_launch<0>(grid, iters, rep, key, h->sink, h->stream); break;
        case 1: mb_launch<1>(grid, iters, rep, key, h->sink, h->stream); break;
        case 2: mb_launch<2>(grid, iters, rep, key, h->sink, h->stream); break;
        case 3: mb_launch<3>(grid, iters, rep, key, h->sink, h->stream); break;
        case 4: mb_launch<4>(grid, iters, rep, key, h->sink, h->stream); break;
        case 5: mb_launch<5>(grid, iters, rep, key, h->sink, h->stream); break;
        case 6: mb_launch<6>(grid, iters, rep, key, h->sink, h->stream); break;
        case 7: mb_launch<7>(grid, iters, rep, key, h->sink, h->stream); break;
        case 8: mb_launch<8>(grid, iters, rep, key, h->sink, h->stream); break;
        case 9: mb_launch<9>(grid, iters, rep, key, h->sink, h->stream); break;
        case 10: mb_launch<10>(grid, iters, rep, key, h->sink, h->stream); break;
        case 11: mb_launch<11>(grid, iters, rep, key, h->sink, h->stream); break;
        case 12: mb_launch<12>(grid, iters, rep, key, h->sink, h->stream); break;
        case 13: mb_launch<13>(grid, iters, rep, key, h->sink, h->stream); break;
        case 14: mb_launch<14>(grid, iters, rep, key, h->sink, h->stream); break;
        case 15: mb_launch<15>(grid, iters, rep, key, h->sink, h->stream); break;
        case 16: mb_launch<16>(grid, iters, rep, key, h->sink, h->stream); break;
        case 17: mb_launch<17>(grid, iters, rep, key, h->sink, h->stream); break;
        case 18: mb_launch<18>(grid, iters, rep, key, h->sink, h->stream); break;
        case 19: mb_launch<19>(grid, iters, rep, key, h->sink, h->stream); break;
        case 20: mb_launch<20>(grid, iters, rep, key, h->sink, h->stream); break;
        default: mb_launch<21>(grid, iters, rep, key, h->sink, h->stream); break;
        }
        PROBE_CUDA(cudaGetLastError());
        PROBE_CUDA(cudaEventRecord(h->ev1, h->stream));
        PROBE_CUDA(cudaEventSynchronize(h->ev1));
        float ms = 0.f;
        PROBE_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double per_thread = (which == MB_PHILOX || which == MB_PHILOX_BM) ? (double)iters
                            : ((which == MB_MIX_FFMA_LOP3 || which >= MB_MIX_FFMA2_LOP3) ? 4.0 * iters : (which == MB_F2F ? 16.0 * iters : 8.0 * iters));
    *ops_per_s = (double)grid * 256.0 * per_thread / ((double)best * 1e-3);
    return 0;
}
