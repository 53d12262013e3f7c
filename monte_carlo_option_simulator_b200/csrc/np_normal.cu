// np_normal.cu -- NumPy's default_rng(seed).standard_normal(...) reproduced on the device, bit for bit.
//
// The reference's pseudo-random front end is three calls of Generator.standard_normal((n_paths, n_steps)) on one PCG64
// generator (engine/monte_carlo.py:301-304, engine/greeks.py:33-41, :458-461): 82 % of its price() time, single-threaded
// (SURVEY.md section 3).  NumPy is a third-party dependency of the reference (numpy 2.3.5 in this image); its published
// algorithm (numpy/random/src/distributions/distributions.c, random_standard_normal) is the 256-layer Ziggurat:
//     r = next_uint64;  idx = r & 0xff;  sign = (r >> 8) & 1;  rabs = (r >> 9) & (2^52 - 1);  x = +-rabs * wi[idx]
//     rabs < ki[idx]                       -> return x                                (98.8 % of the draws: ONE output)
//     idx == 0 (tail)                      -> loop: xx = -log1p(-U) / R, yy = -log1p(-U'); accept when 2 yy > xx^2
//     else (wedge), one more output U      -> return x when (fi[idx-1] - fi[idx]) U + fi[idx] < exp(-x^2 / 2), else retry
// so a normal consumes a DATA-DEPENDENT number of generator outputs and the stream looks sequential.  It is not:
//   * every output index i is the possible start of an "attempt" whose length c_i and result depend only on the outputs
//     at i, i+1, ... -- computable independently for every i (the LCG jumps ahead in O(log i));
//   * the attempts actually taken are the chain 0 -> c_0 -> c_0 + c_{c_0} -> ...; two chains started at different
//     indices MERGE at the first index both visit, and because 98.8 % of the attempts have length 1 they merge
//     within a few outputs.
// Hence: cut the output stream into chunks of 128; (1) every chunk walks the chain speculatively from its FIRST index
// and records where that chain leaves the chunk; (2) every chunk walks again from the exit of its predecessor's
// speculative chain -- which is the true chain's exit unless the two failed to merge inside one chunk (probability
// ~0.02^40; detected exactly: the walk's own exit must equal the speculative one, else the host repeats step 2 with
// the corrected exits until nothing changes) -- and counts the normals it emits; (3) an exclusive scan of the counts
// places every chunk's normals; (4) a last walk writes them.  Results are NumPy's doubles bit for bit, including the
// tail draws: log1p follows glibc's s_log1p.c as x86-64 hosts with FMA run it (the FMA-contracted multiarch build;
// checked bitwise against libm on 1e8 arguments, oracle/svj_oracle.c oracle_glibc_log1p_fma); exp only feeds the wedge
// comparison (a last-bit difference flips the decision with probability ~1e-14 per wedge draw).
#include <math_constants.h>

#include "pcg64.cuh"

namespace b200mc {

namespace zig_dev {
#define NP_ZIG_STORAGE static __device__ const
#include "np_ziggurat_tables.inc"
#undef NP_ZIG_STORAGE
} // namespace zig_dev
namespace zig_host {
#define NP_ZIG_STORAGE static const
#include "np_ziggurat_tables.inc"
#undef NP_ZIG_STORAGE
} // namespace zig_host

constexpr int ZG_CHUNK = 128;        // generator outputs per chunk (= per thread)
constexpr int ZG_THREADS = 128;
constexpr double ZIG_R = 3.6541528853610087963519472518;
constexpr double ZIG_INV_R = 0.27366123732975827203338247596;

struct ZigTables {
    unsigned long long ki[256];
    double wi[256], fi[256];
};

__device__ __forceinline__ void load_tables(ZigTables &t)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        t.ki[i] = zig_dev::NP_ZIG_KI[i];
        t.wi[i] = __longlong_as_double((long long)zig_dev::NP_ZIG_WI[i]);
        t.fi[i] = __longlong_as_double((long long)zig_dev::NP_ZIG_FI[i]);
    }
    __syncthreads();
}

// log1p(x) for x in (-1, 0]: glibc 2.28+ sysdeps/ieee754/dbl-64/s_log1p.c in the operation order AND the fused
// multiply-adds of its x86-64 FMA build (what numpy's npy_log1p resolves to on hosts with FMA).  Every product / sum
// is rounded separately (__dmul_rn / __dadd_rn keep nvcc from contracting) except where the host build fuses (__fma_rn).
__device__ __forceinline__ double np_log1p(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    double f = 0.0, c = 0.0;
    int hu = 0, k = 1;
    const int hx = __double2hiint(x), ax = hx & 0x7fffffff;
    if (hx < 0x3FDA827A) {                                   // x < 0.41422
        if (ax >= 0x3ff00000) return x == -1.0 ? -CUDART_INF : CUDART_NAN;
        if (ax < 0x3e200000) {                               // |x| < 2^-29
            if (ax < 0x3c900000) return x;
            return __fma_rn(-__dmul_rn(x, x), 0.5, x);
        }
        if (hx > 0 || hx <= (int)0xbfd2bec3) { k = 0; f = x; hu = 1; }      // -0.2929 < x < 0.41422
    }
    if (k != 0) {
        double u = __dadd_rn(1.0, x);
        hu = __double2hiint(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? __dsub_rn(1.0, __dsub_rn(u, x)) : __dsub_rn(x, __dsub_rn(u, 1.0));
        c = __ddiv_rn(c, u);
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = __hiloint2double(hu | 0x3ff00000, __double2loint(u));
        } else {
            k += 1;
            u = __hiloint2double(hu | 0x3fe00000, __double2loint(u));
            hu = (0x00100000 - hu) >> 2;
        }
        f = __dsub_rn(u, 1.0);
    }
    const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
    const double dk = (double)k;
    if (hu == 0) {                                           // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = __fma_rn(dk, ln2_lo, c);
            return __fma_rn(dk, ln2_hi, c);
        }
        const double R = __dmul_rn(hfsq, __fma_rn(-0.66666666666666666, f, 1.0));
        if (k == 0) return __dsub_rn(f, R);
        return __fma_rn(dk, ln2_hi, -__dsub_rn(__dsub_rn(R, __fma_rn(dk, ln2_lo, c)), f));
    }
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s);
    const double z2 = __dmul_rn(z, z), R2 = __fma_rn(z, Lp3, Lp2), z4 = __dmul_rn(z2, z2), R3 = __fma_rn(z, Lp5, Lp4),
                 z6 = __dmul_rn(z4, z2), R4 = __fma_rn(z, Lp7, Lp6);
    const double R = __fma_rn(z6, R4, __fma_rn(z4, R3, __fma_rn(z, Lp1, __dmul_rn(z2, R2))));
    const double sp = __dmul_rn(s, __dadd_rn(hfsq, R));
    if (k == 0) return __dsub_rn(f, __dsub_rn(hfsq, sp));
    return __fma_rn(dk, ln2_hi, -__dsub_rn(__dsub_rn(hfsq, __dadd_rn(sp, __fma_rn(dk, ln2_lo, c))), f));
}

// One attempt of random_standard_normal starting at the generator's current position.  Returns the number of outputs
// consumed; emit = whether a normal came out (a rejected wedge draw returns without one: the NEXT attempt starts right
// after it, which is all the reference's `for (;;)` does).
__device__ __forceinline__ int zig_attempt(U128 &s, const U128 inc, const ZigTables &t, bool &emit, double &val)
{
    unsigned long long r = pcg64_next(s, inc);
    const int idx = (int)(r & 0xffull);
    r >>= 8;
    const bool neg = (r & 1ull) != 0ull;
    const unsigned long long rabs = (r >> 1) & 0x000fffffffffffffull;
    double x = __dmul_rn((double)rabs, t.wi[idx]);
    if (neg) x = -x;
    val = x;
    emit = true;
    if (rabs < t.ki[idx]) return 1;
    if (idx == 0) {                                                               // the tail beyond R
        int used = 1;
        for (;;) {
            const double xx = __dmul_rn(-ZIG_INV_R, np_log1p(-pcg64_double(pcg64_next(s, inc))));
            const double yy = -np_log1p(-pcg64_double(pcg64_next(s, inc)));
            used += 2;
            if (__dadd_rn(yy, yy) > __dmul_rn(xx, xx)) {
                const double m = __dadd_rn(ZIG_R, xx);
                val = ((rabs >> 8) & 1ull) ? -m : m;
                return used;
            }
        }
    }
    const double u = pcg64_double(pcg64_next(s, inc));                            // a wedge
    const double lhs = __dadd_rn(__dmul_rn(__dsub_rn(t.fi[idx - 1], t.fi[idx]), u), t.fi[idx]);
    emit = lhs < exp(__dmul_rn(__dmul_rn(-0.5, x), x));
    return 2;
}

struct ZigArgs {
    U128 state, inc;               // generator state BEFORE output `first`
    unsigned long long first;      // outputs already consumed by earlier calls on this generator
    long long n_chunks;
    long long n_normals;
};

// state of the generator at the first output of every chunk (one O(log) jump per chunk, reused by the three walks)
__global__ void __launch_bounds__(ZG_THREADS) k_zig_states(const __grid_constant__ ZigArgs a, U128 *__restrict__ states)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < a.n_chunks) states[k] = pcg64_advance(a.state, a.inc, a.first + (unsigned long long)k * ZG_CHUNK);
}

// MODE 0: speculative walk from the chunk's first output            -> exit_out
// MODE 1: walk from the predecessor's exit (exit_in)                -> exit_out, count, *changed |= (exit_out != old)
// MODE 2: walk from the predecessor's exit and write the normals at offs[k] + j
template <int MODE>
__global__ void __launch_bounds__(ZG_THREADS)
k_zig_walk(const __grid_constant__ ZigArgs a, const U128 *__restrict__ states, const int *exit_in,
           int *exit_out /* MODE 1: the same array as exit_in */, int *__restrict__ count, const long long *__restrict__ offs, int *changed,
           double *__restrict__ out, unsigned long long *consumed)
{
    __shared__ ZigTables t;
    load_tables(t);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_chunks) return;
    U128 s = states[k];
    long long pos = 0;                                   // relative to the chunk's first output
    if (MODE != 0 && k > 0) {
        const int e = exit_in[k - 1];
        for (; pos < e; ++pos) pcg64_next(s, a.inc);     // skip the outputs the predecessor's last attempt consumed
    }
    int cnt = 0;
    long long j = MODE == 2 ? offs[k] : 0;
    while (pos < ZG_CHUNK) {
        bool emit;
        double val;
        pos += zig_attempt(s, a.inc, t, emit, val);
        if (emit) {
            if (MODE == 2) {
                if (j < a.n_normals) out[j] = val;
                if (j == a.n_normals - 1) *consumed = (unsigned long long)k * ZG_CHUNK + (unsigned long long)pos;
                ++j;
            }
            ++cnt;
        }
    }
    const int ex = (int)(pos - ZG_CHUNK);
    if (MODE == 0) exit_out[k] = ex;
    if (MODE == 1) {
        if (exit_out[k] != ex) { exit_out[k] = ex; atomicOr(changed, 1); }
        count[k] = cnt;
    }
}

// ---- exclusive scan of the per-chunk counts (int -> long long), three small kernels ------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long *warp_sums, long long &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < SC_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < SC_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    total = warp_sums[SC_THREADS / 32 - 1];
    const long long before = warp > 0 ? warp_sums[warp - 1] : 0;
    return before + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_tiles(const int *__restrict__ cnt, long long n, long long *__restrict__ offs,
                                                           long long *__restrict__ tile_sums)
{
    __shared__ long long ws[SC_THREADS / 32];
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
    int c[SC_ITEMS];
    long long mine = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) { c[i] = base + i < n ? cnt[base + i] : 0; mine += c[i]; }
    long long total;
    long long run = block_exclusive_scan(mine, ws, total);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) { if (base + i < n) offs[base + i] = run; run += c[i]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place; total -> *grand
__global__ void __launch_bounds__(SC_THREADS) k_scan_sums(long long *__restrict__ tile_sums, long long n_tiles, long long *grand)
{
    __shared__ long long ws[SC_THREADS / 32];
    long long carry = 0;
    for (long long b0 = 0; b0 < n_tiles; b0 += SC_THREADS) {
        const long long i = b0 + threadIdx.x;
        const long long v = i < n_tiles ? tile_sums[i] : 0;
        long long total;
        const long long ex = block_exclusive_scan(v, ws, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_add(long long *__restrict__ offs, long long n, const long long *__restrict__ tile_sums)
{
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
    const long long add = tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i)
        if (base + i < n) offs[base + i] += add;
}

struct ZigResult {
    long long total;                     // normals the examined outputs yield
    unsigned long long consumed;         // outputs consumed by the first n_normals normals
    int changed;
    int pad_;
};

// n normals of default_rng's stream starting at output index `first` -> out_dev; *consumed_out = generator outputs used.
static int np_normal_fill(b200mc_handle *h, const uint64_t st[4], uint64_t first, int64_t n, double *out_dev,
                          uint64_t *consumed_out)
{
    for (int attempt = 0; attempt < 4; ++attempt) {
        // outputs to examine: 1.0215 per normal on average; 3 % + slack, doubled on the (never observed) retry
        const int64_t m = n + (n >> (5 - attempt > 1 ? 5 - attempt : 1)) + 8192;
        const int64_t nch = (m + ZG_CHUNK - 1) / ZG_CHUNK;
        const int64_t ntile = (nch + SC_TILE - 1) / SC_TILE;
        // scratch: [states nch*16][exit nch*4][count nch*4][offs nch*8][tile sums][result]
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_st = carve((size_t)nch * 16), o_ex = carve((size_t)nch * 4), o_cn = carve((size_t)nch * 4),
                     o_of = carve((size_t)nch * 8), o_ts = carve((size_t)ntile * 8), o_rs = carve(sizeof(ZigResult));
        B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, off));
        char *sc = (char *)h->d_scratch;
        U128 *states = (U128 *)(sc + o_st);
        int *ex = (int *)(sc + o_ex), *cn = (int *)(sc + o_cn);
        long long *offs = (long long *)(sc + o_of), *tsum = (long long *)(sc + o_ts);
        ZigResult *res = (ZigResult *)(sc + o_rs);
        ZigArgs a;
        a.state = {st[0], st[1]};
        a.inc = {st[2], st[3]};
        a.first = first;
        a.n_chunks = nch;
        a.n_normals = n;
        const unsigned grid = (unsigned)((nch + ZG_THREADS - 1) / ZG_THREADS);
        B200MC_CUDA(h, cudaMemsetAsync(res, 0, sizeof(ZigResult), h->stream));
        k_zig_states<<<grid, ZG_THREADS, 0, h->stream>>>(a, states);
        k_zig_walk<0><<<grid, ZG_THREADS, 0, h->stream>>>(a, states, nullptr, ex, nullptr, nullptr, nullptr, nullptr, nullptr);
        h->launches += 2;
        ZigResult r;
        for (int64_t round = 0; round <= nch; ++round) {
            B200MC_CUDA(h, cudaMemsetAsync(&res->changed, 0, sizeof(int), h->stream));
            k_zig_walk<1><<<grid, ZG_THREADS, 0, h->stream>>>(a, states, ex, ex, cn, nullptr, &res->changed, nullptr, nullptr);
            k_scan_tiles<<<(unsigned)ntile, SC_THREADS, 0, h->stream>>>(cn, nch, offs, tsum);
            k_scan_sums<<<1, SC_THREADS, 0, h->stream>>>(tsum, ntile, &res->total);
            k_scan_add<<<(unsigned)ntile, SC_THREADS, 0, h->stream>>>(offs, nch, tsum);
            k_zig_walk<2><<<grid, ZG_THREADS, 0, h->stream>>>(a, states, ex, nullptr, nullptr, offs, nullptr, out_dev, &res->consumed);
            B200MC_CUDA(h, cudaGetLastError());
            h->launches += 5;
            B200MC_CUDA(h, cudaMemcpyAsync(&r, res, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
            B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
            if (!r.changed) break;      // every chunk entered where its predecessor's chain really left: the walk was the true one
        }
        if (r.total >= n) {
            if (consumed_out) *consumed_out = r.consumed;
            return 0;
        }
    }
    return fail(h, B200MC_ECUDA, "standard_normal: the examined generator outputs did not yield the requested normals");
}

} // namespace b200mc
using namespace b200mc;

// kind 0: n doubles of Generator.random() (one output each); kind 1: n doubles of Generator.standard_normal() -- both
// starting at output index first_raw of the PCG64 generator whose (state, inc) the caller read from NumPy
// (default_rng(seed).bit_generator.state).  out is a DEVICE pointer (on_device) or a host array.
extern "C" int b200mc_numpy_fill(b200mc_handle *h, const uint64_t state[4], uint64_t first_raw, int64_t n, int kind,
                                 int on_device, double *out, uint64_t *raws_consumed)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!state || !out || n <= 0) return fail(h, B200MC_EINVAL, "state / out must be given and n positive");
    if (kind != B200MC_NUMPY_RANDOM && kind != B200MC_NUMPY_STANDARD_NORMAL) return fail(h, B200MC_EINVAL, "unknown kind");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    double *dst = out;
    if (!on_device) {
        B200MC_TRY(ensure(h, &h->d_stage, &h->stage_bytes, (size_t)n * 8));
        dst = (double *)h->d_stage;
    }
    if (kind == B200MC_NUMPY_RANDOM) {
        B200MC_TRY(pcg64_uniform_async(h, state, first_raw, n, 64, dst));
        if (raws_consumed) *raws_consumed = (uint64_t)n;
    } else {
        B200MC_TRY(np_normal_fill(h, state, first_raw, n, dst, raws_consumed));
    }
    if (!on_device) {
        B200MC_CUDA(h, cudaMemcpyAsync(out, dst, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}

// The Ziggurat tables the kernels use (numpy's ki_double / wi_double / fi_double), for inspection and tests.  Host only.
extern "C" int b200mc_numpy_ziggurat_tables(uint64_t ki[256], double wi[256], double fi[256])
{
    if (!ki || !wi || !fi) return B200MC_EINVAL;
    for (int i = 0; i < 256; ++i) {
        ki[i] = zig_host::NP_ZIG_KI[i];
        memcpy(&wi[i], &zig_host::NP_ZIG_WI[i], 8);
        memcpy(&fi[i], &zig_host::NP_ZIG_FI[i], 8);
    }
    return 0;
}
