// qmc.cu -- a WORKING quasi-Monte Carlo front end, offered beside the reference's own.
//
// The reference's use_sobol=True path (engine/monte_carlo.py:61-183) is degenerate: its "Brownian bridge" places the
// endpoint with conditional variance 0, so W_T == 0 on every path.  That behaviour stays
// reproducible through rng="reference" (host front end).  This file is what that path was meant to be:
//   1. scrambled Sobol points on the device, BITWISE those of scipy.stats.qmc.Sobol(d, scramble=True, seed): the caller
//      passes the scrambled direction numbers sv[d][bits] and the digital shift[d] of that engine (SciPy computes them in
//      O(d bits); the library embeds no direction-number table) and point n is shift ^ XOR_{b in gray(n)} sv[.][b];
//   2. u = clip(x 2^-bits, 1e-10, 1 - 1e-10), z = normcdfinv(u)          (monte_carlo.py:80-84, same clip)
//   3. a correct Brownian bridge for the two Brownian drivers: dimension 0 places W_T ~ N(0, n_steps), the following
//      dimensions the interval midpoints breadth first, W_m | W_l, W_r ~ N(((r-m) W_l + (m-l) W_r)/(r-l),
//      (m-l)(r-m)/(r-l)); the increments W_{s+1} - W_s are the step normals (unit variance, unit time steps);
//   4. the step normals go, device resident, through the fp64 given-normals kernel (given_normals.cu) -- the same
//      recurrence as everything else -- and the terminal values are reduced to b200mc_sums on the device.
// Dimension layout of the point set: [0, s) Z1 in bridge order, [s, 2s) Z2 in bridge order, [2s, 3s) jump sizes in time
// order, [3s, 4s) jump uniforms in time order (s = n_steps); only the blocks the parameters need are generated.
#include <algorithm>
#include <vector>

#include "prep.cuh"

extern "C" int b200mc_simulate_given_normals_dev(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                                 int64_t n_paths, int32_t n_steps, const double *Z1, const double *Z2,
                                                 const double *Z_jump, const double *Z_jump_size, int record_paths,
                                                 double *S_final, double *v_final, double *all_paths);

namespace b200mc {

constexpr int QMC_PT_MAX = 32;      // paths per CTA tile: 32, 16 or 8 (the largest whose tile leaves >= 4 CTAs per SM)
constexpr int QMC_THREADS = 256;

using BridgeNode = b200mc_bridge_node;   // W[t] = W[l] + ((W[r] - W[l]) * a) / b + sd * z[dim]   (index 0 = the known W_0 = 0)

// The built-in, CORRECT bridge over the grid 0..n in construction order: the endpoint first (W_n ~ N(0, n)), then the
// midpoints of the open intervals breadth first: W_m | W_l, W_r ~ N(W_l + (W_r - W_l)(m-l)/(r-l), (m-l)(r-m)/(r-l)).
static std::vector<BridgeNode> bridge_table(int n)
{
    std::vector<BridgeNode> nodes;
    nodes.reserve((size_t)n);
    nodes.push_back({n, 0, 0, 0, 0.0, 1.0, sqrt((double)n)});
    std::vector<std::pair<int, int>> cur{{0, n}}, next;
    while (!cur.empty()) {
        next.clear();
        for (auto [l, r] : cur) {
            if (r - l <= 1) continue;
            const int m = (l + r) / 2;
            nodes.push_back({m, l, r, (int32_t)nodes.size(), (double)(m - l), (double)(r - l),
                             sqrt((double)(m - l) * (double)(r - m) / (double)(r - l))});
            next.push_back({l, m});
            next.push_back({m, r});
        }
        cur.swap(next);
    }
    return nodes;
}

// Nodes in construction order -> nodes grouped by dependency level (a node needs W[l] and W[r] of earlier nodes), so that
// a level can be built by all threads at once.  Returns false when the table is not a valid construction.
static bool schedule_nodes(const BridgeNode *in, int n_nodes, int n_steps, std::vector<BridgeNode> &out,
                           std::vector<int32_t> &level_start)
{
    std::vector<int> level_of_index((size_t)n_steps + 1, -1), lvl((size_t)n_nodes, 0);
    level_of_index[0] = 0;                                   // W_0 = 0 is known from the start
    int max_level = 0;
    for (int k = 0; k < n_nodes; ++k) {
        const BridgeNode &nd = in[k];
        if (nd.t < 1 || nd.t > n_steps || nd.l < 0 || nd.l > n_steps || nd.r < 0 || nd.r > n_steps || nd.dim < 0 ||
            nd.dim >= n_steps || !(nd.b != 0.0) || level_of_index[nd.t] >= 0 || level_of_index[nd.l] < 0 ||
            level_of_index[nd.r] < 0)
            return false;
        lvl[k] = 1 + std::max(level_of_index[nd.l], level_of_index[nd.r]);
        level_of_index[nd.t] = lvl[k];
        max_level = std::max(max_level, lvl[k]);
    }
    out.clear();
    level_start.clear();
    for (int L = 1; L <= max_level; ++L) {
        level_start.push_back((int32_t)out.size());
        for (int k = 0; k < n_nodes; ++k)
            if (lvl[k] == L) out.push_back(in[k]);
    }
    level_start.push_back((int32_t)out.size());
    return true;
}

struct QmcArgs {
    uint64_t path0;
    int64_t n_paths;
    int32_t n_steps, bits, pitch, n_levels, pt, n_set;    // n_set: every W[1..n_set] is written by the table
    double sign;                    // +1, or -1 for the antithetic pass (normals negated, uniforms kept)
    double scale;                   // 2^-bits
};

__device__ __forceinline__ uint32_t sobol_point(const uint32_t *__restrict__ sv, uint32_t shift, int bits, uint64_t n)
{
    uint64_t g = n ^ (n >> 1);
    uint32_t x = shift;
    for (int b = 0; b < bits && g; ++b, g >>= 1)
        if (g & 1) x ^= sv[b];
    return x;
}

// which: 0 = bridged normals (dims [dim0, dim0 + n_steps) in bridge order), 1 = plain normals in time order,
// 2 = uniforms in time order.  out: [n_paths][n_steps] float64.
__global__ void __launch_bounds__(QMC_THREADS)
k_qmc_block(const __grid_constant__ QmcArgs a, int which, const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
            const BridgeNode *__restrict__ nodes, const int32_t *__restrict__ level_start, double *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *zb = reinterpret_cast<double *>(smem_raw);              // [pt][pitch]  draws, dimension-major per path
    const int tid = threadIdx.x;
    const int s = a.n_steps, PT = a.pt;
    double *W = zb + (size_t)PT * a.pitch;                          // [pt][pitch]  bridge values (which == 0)
    for (int64_t tile = blockIdx.x; tile * PT < a.n_paths; tile += gridDim.x) {
        const int64_t base = tile * PT;
        const int np = (int)min((int64_t)PT, a.n_paths - base);
        // ---- draws: item = (dimension j, chunk of consecutive paths).  The first point of a chunk is evaluated from
        // the gray code of its index, the following ones by the gray-code increment x_{n+1} = x_n ^ sv[j][ctz(n + 1)]:
        // one load and one XOR per point instead of a loop over the set bits.  Lanes run over j: stride-1 smem stores.
        {
            const int nsub = max(1, min(PT, QMC_THREADS / s));             // chunks per tile (more when s is small)
            const int len = (PT + nsub - 1) / nsub;
            for (int it = tid; it < s * nsub; it += QMC_THREADS) {
                const int j = it % s, sub = it / s;
                const int p0 = sub * len, p1 = min(np, p0 + len);
                if (p0 < p1) {
                    const uint32_t *svj = sv + (size_t)j * a.bits;
                    uint64_t n = a.path0 + (uint64_t)(base + p0);
                    uint32_t x = sobol_point(svj, shift[j], a.bits, n);
                    for (int p = p0; p < p1; ++p) {
                        double u = (double)x * a.scale;
                        u = fmin(fmax(u, 1e-10), 1.0 - 1e-10);                               // monte_carlo.py:83
                        zb[(size_t)p * a.pitch + j] = which == 2 ? u : a.sign * normcdfinv(u);   // :84
                        ++n;
                        const int c = __ffsll((long long)n) - 1;                             // lowest set bit of n
                        if (c < a.bits) x ^= svj[c];
                    }
                }
            }
        }
        __syncthreads();
        if (which == 0) {
            // ---- bridge, level by level: item = (node of the level, path) ---------------------------------------
            for (int lv = 0; lv < a.n_levels; ++lv) {
                const int k0 = level_start[lv], k1 = level_start[lv + 1];
                for (int it = tid; it < (k1 - k0) * PT; it += QMC_THREADS) {
                    const int k = k0 + it / PT, p = it % PT;
                    if (p < np) {
                        const BridgeNode nd = nodes[k];
                        double *w = W + (size_t)p * a.pitch;
                        const double wl = nd.l == 0 ? 0.0 : w[nd.l], wr = nd.r == 0 ? 0.0 : w[nd.r];
                        // the reference's operation order, monte_carlo.py:128-133 (no FMA contraction)
                        const double mean = __dadd_rn(wl, __ddiv_rn(__dmul_rn(__dadd_rn(wr, -wl), nd.a), nd.b));
                        w[nd.t] = __dadd_rn(mean, __dmul_rn(nd.sd, zb[(size_t)p * a.pitch + nd.dim]));
                    }
                }
                __syncthreads();
            }
        }
        // ---- write [path][step], step fastest -------------------------------------------------------------------
        for (int it = tid; it < np * s; it += QMC_THREADS) {
            const int p = it / s, t = it % s;
            double v;
            if (which == 0) {
                const double *w = W + (size_t)p * a.pitch;
                v = (t + 1 <= a.n_set ? w[t + 1] : 0.0) - (t == 0 ? 0.0 : w[t]);
            } else {
                v = zb[(size_t)p * a.pitch + t];
            }
            out[(size_t)(base + p) * s + t] = v;
        }
        __syncthreads();
    }
}

// Terminal values -> b200mc_sums (price part) for n_strikes strikes; one launch, deterministic finish.
constexpr int TS_NV = 8;
__global__ void __launch_bounds__(256)
k_terminal_sums(const double *__restrict__ S, const double *__restrict__ A, int64_t n, double K, int is_call, int anti,
                double *partials, unsigned int *counter, double *out)
{
    __shared__ double smem[(256 / 32) * TS_NV];
    double v[TS_NV];
#pragma unroll
    for (int i = 0; i < TS_NV; ++i) v[i] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sa = S[i];
        const double da = is_call ? fmax(sa - K, 0.0) : fmax(K - sa, 0.0);
        double db = 0.0, s_avg = sa, pay = da;
        if (anti) {
            const double sb = A[i];
            db = is_call ? fmax(sb - K, 0.0) : fmax(K - sb, 0.0);
            s_avg = 0.5 * (sa + sb);
            pay = 0.5 * (da + db);
        }
        v[0] += da; v[1] += db; v[2] += da * da; v[3] += db * db; v[4] += da * db;
        v[5] += s_avg; v[6] += s_avg * s_avg; v[7] += pay * s_avg;
    }
    block_finish<TS_NV>(v, smem, partials, counter, out);
}

static int upload(b200mc_handle *h, void *dst, const void *src, size_t bytes)
{
    B200MC_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

struct QmcPlan {
    int n_steps, bits, pitch, n_levels, n_blocks, pt;      // n_blocks: 1 (Z1), 2 (+Z2) or 4 (+ jump sizes and uniforms)
    size_t smem;
    uint32_t *d_sv, *d_shift;
    BridgeNode *d_nodes;
    int32_t *d_levels;
};

// Tables -> device (inside d_scratch).  Returns the bytes used at the front of the scratch buffer.
static int qmc_prepare(b200mc_handle *h, int32_t n_steps, const uint32_t *sv, const uint32_t *shift, int32_t n_dims,
                       int32_t bits, int n_blocks, QmcPlan &pl, size_t extra_bytes, char **extra,
                       const BridgeNode *custom = nullptr, int32_t n_custom = 0)
{
    if (!sv || !shift) return fail(h, B200MC_EINVAL, "sv / shift is NULL");
    if (bits < 1 || bits > 32) return fail(h, B200MC_EINVAL, "bits must be in [1, 32]");
    if (n_dims < n_blocks * n_steps) return fail(h, B200MC_EINVAL, "the point set has fewer dimensions than the run needs");
    pl.n_steps = n_steps; pl.bits = bits; pl.n_blocks = n_blocks;
    pl.pitch = (n_steps + 1) | 1;                       // odd pitch (in doubles): conflict-free with lanes over paths
    pl.pt = QMC_PT_MAX;                                 // 128 KB tiles (1 CTA per SM) were latency bound: 0.62 ms per 64k x 250
    while (pl.pt > 8 && (size_t)2 * pl.pt * pl.pitch * sizeof(double) > 56 * 1024) pl.pt >>= 1;
    while (pl.pt > 1 && (size_t)2 * pl.pt * pl.pitch * sizeof(double) > (size_t)h->smem_optin - 1024) pl.pt >>= 1;   // very long paths
    pl.smem = (size_t)2 * pl.pt * pl.pitch * sizeof(double);
    if (pl.smem > (size_t)h->smem_optin - 1024) return fail(h, B200MC_EINVAL, "too many steps for the bridge tile");
    std::vector<BridgeNode> built, nodes;
    std::vector<int32_t> levels;
    if (!custom) built = bridge_table(n_steps);
    else if (n_custom != n_steps) return fail(h, B200MC_EINVAL, "a bridge table must place every one of the n_steps points");
    if (!schedule_nodes(custom ? custom : built.data(), n_steps, n_steps, nodes, levels))
        return fail(h, B200MC_EINVAL, "the bridge table is not a valid construction (indices, order or b == 0)");
    pl.n_levels = (int)levels.size() - 1;
    const size_t nd = (size_t)n_blocks * n_steps;
    const size_t b_sv = nd * bits * 4, b_sh = nd * 4, b_nodes = nodes.size() * sizeof(BridgeNode), b_lv = levels.size() * 4;
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t total = up16(b_sv) + up16(b_sh) + up16(b_nodes) + up16(b_lv);
    B200MC_TRY(ensure(h, &h->d_scratch, &h->scratch_bytes, total + extra_bytes + 64));
    B200MC_TRY(ensure(h, &h->h_pinned, &h->pinned_bytes, total, true));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    char *hp = (char *)h->h_pinned, *dp = (char *)h->d_scratch;
    size_t off = 0;
    memcpy(hp + off, sv, b_sv);              pl.d_sv = (uint32_t *)(dp + off);      off += up16(b_sv);
    memcpy(hp + off, shift, b_sh);           pl.d_shift = (uint32_t *)(dp + off);   off += up16(b_sh);
    memcpy(hp + off, nodes.data(), b_nodes); pl.d_nodes = (BridgeNode *)(dp + off); off += up16(b_nodes);
    memcpy(hp + off, levels.data(), b_lv);   pl.d_levels = (int32_t *)(dp + off);   off += up16(b_lv);
    B200MC_TRY(upload(h, dp, hp, total));
    *extra = dp + total;
    B200MC_CUDA(h, cudaFuncSetAttribute((const void *)k_qmc_block, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        h->smem_optin - 1024));
    return 0;
}

// Block b of the point set (0 Z1, 1 Z2: bridged; 2 jump sizes: plain normals; 3 jump uniforms) for paths
// [path0, path0 + n) -> out[n][n_steps] (device).
static int qmc_block(b200mc_handle *h, const QmcPlan &pl, int b, uint64_t path0, int64_t n, double sign, double *out)
{
    QmcArgs a;
    a.path0 = path0; a.n_paths = n; a.n_steps = pl.n_steps; a.bits = pl.bits; a.pitch = pl.pitch; a.n_levels = pl.n_levels;
    a.pt = pl.pt; a.n_set = pl.n_steps;
    a.sign = sign;
    a.scale = ldexp(1.0, -pl.bits);
    const int which = b <= 1 ? 0 : (b == 2 ? 1 : 2);
    const int64_t tiles = (n + pl.pt - 1) / pl.pt;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)h->sm_count * 16);
    k_qmc_block<<<grid, QMC_THREADS, pl.smem, h->stream>>>(a, which, pl.d_sv + (size_t)b * pl.n_steps * pl.bits,
                                                           pl.d_shift + (size_t)b * pl.n_steps, pl.d_nodes, pl.d_levels, out);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

static int blocks_needed(const b200mc_svj_params *p, double T, int32_t n_steps)
{
    if (p->lambda_j * (T / n_steps) > 0.0) return 4;
    return p->xi != 0.0 ? 2 : 1;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_qmc_normals(b200mc_handle *h, int64_t n_paths, uint64_t path_offset, int32_t n_steps,
                                  const uint32_t *sv, const uint32_t *shift, int32_t n_dims, int32_t bits, int which,
                                  double *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (n_paths <= 0 || n_steps <= 0) return fail(h, B200MC_EINVAL, "n_paths and n_steps must be positive");
    if (which < 0 || which > 3) return fail(h, B200MC_EINVAL, "which must be one of B200MC_Z1 .. B200MC_ZJUMP_SIZE");
    if (!out) return fail(h, B200MC_EINVAL, "out is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    // ABI order of `which` (B200MC_Z1, Z2, ZJUMP_U, ZJUMP_SIZE) -> block of the point set (Z1, Z2, sizes, uniforms)
    const int block = which == B200MC_ZJUMP_U ? 3 : (which == B200MC_ZJUMP_SIZE ? 2 : which);
    QmcPlan pl;
    char *extra = nullptr;
    const size_t bytes = (size_t)n_paths * n_steps * 8;
    B200MC_TRY(qmc_prepare(h, n_steps, sv, shift, n_dims, bits, block + 1, pl, bytes, &extra));
    B200MC_TRY(qmc_block(h, pl, block, path_offset, n_paths, 1.0, (double *)extra));
    B200MC_CUDA(h, cudaMemcpyAsync(out, extra, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int b200mc_price_european_qmc(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T,
                                         int32_t n_steps, int64_t n_paths, uint64_t path_offset, const uint32_t *sv,
                                         const uint32_t *shift, int32_t n_dims, int32_t bits, const double *strikes,
                                         int32_t n_strikes, int is_call, uint32_t flags, b200mc_sums *out)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!p) return fail(h, B200MC_EINVAL, "params is NULL");
    if (!out || !strikes || n_strikes <= 0) return fail(h, B200MC_EINVAL, "strikes / out must hold at least one entry");
    if (n_paths <= 0 || n_steps <= 0) return fail(h, B200MC_EINVAL, "n_paths and n_steps must be positive");
    if (!(T > 0.0) || !isfinite(T) || !isfinite(S0)) return fail(h, B200MC_EINVAL, "T must be positive and finite, S0 finite");
    if (flags & B200MC_GREEKS) return fail(h, B200MC_EINVAL, "the quasi-Monte Carlo entry point has no Greek sums");
    if (path_offset + (uint64_t)n_paths > (1ull << (bits < 1 ? 1 : bits)))
        return fail(h, B200MC_EINVAL, "the Sobol sequence holds 2^bits points");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const bool anti = flags & B200MC_ANTITHETIC;
    const int nb = blocks_needed(p, T, n_steps);

    // chunks of paths so that the step normals stay below ~2 GiB per block
    int64_t chunk = std::max<int64_t>(QMC_PT_MAX, ((int64_t)1 << 28) / n_steps);
    chunk = std::min(chunk, n_paths);
    const size_t zbytes = (size_t)chunk * n_steps * 8;
    const int grid_r = (int)std::min<int64_t>((chunk + 255) / 256, (int64_t)h->sm_count * 4);
    // extra scratch: [Z x nb][S][A][v][partials][sums n_strikes x 8]
    const size_t extra_bytes = zbytes * nb + (size_t)chunk * 8 * 3 + (size_t)grid_r * TS_NV * 8 + (size_t)n_strikes * TS_NV * 8;
    QmcPlan pl;
    char *extra = nullptr;
    B200MC_TRY(qmc_prepare(h, n_steps, sv, shift, n_dims, bits, nb, pl, extra_bytes, &extra));
    double *Z[4];
    for (int b = 0; b < 4; ++b) Z[b] = (double *)(extra + zbytes * (b < nb ? b : 0));
    double *dS = (double *)(extra + zbytes * nb), *dA = dS + chunk, *dV = dA + chunk, *dPart = dV + chunk,
           *dSums = dPart + (size_t)grid_r * TS_NV;
    std::vector<double> acc((size_t)n_strikes * TS_NV, 0.0), tmp((size_t)n_strikes * TS_NV);

    for (int64_t done = 0; done < n_paths; done += chunk) {
        const int64_t n = std::min(chunk, n_paths - done);
        for (int pass = 0; pass < (anti ? 2 : 1); ++pass) {
            const double sign = pass ? -1.0 : 1.0;                      // twin: -Z1, -Z2, U, -Zjs (monte_carlo.py:323)
            for (int b = 0; b < nb; ++b) B200MC_TRY(qmc_block(h, pl, b, path_offset + (uint64_t)done, n, sign, Z[b]));
            B200MC_TRY(b200mc_simulate_given_normals_dev(h, p, S0, T, n, n_steps, Z[0], Z[1], Z[3], Z[2], 0,
                                                         pass ? dA : dS, dV, nullptr));
        }
        for (int k = 0; k < n_strikes; ++k) {
            k_terminal_sums<<<grid_r, 256, 0, h->stream>>>(dS, dA, n, strikes[k], is_call ? 1 : 0, anti ? 1 : 0, dPart,
                                                           h->d_counter, dSums + (size_t)k * TS_NV);
            B200MC_CUDA(h, cudaGetLastError());
            h->launches += 1;
        }
        B200MC_CUDA(h, cudaMemcpyAsync(tmp.data(), dSums, tmp.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
        for (size_t i = 0; i < acc.size(); ++i) acc[i] += tmp[i];
    }
    memset(out, 0, (size_t)n_strikes * sizeof(b200mc_sums));
    for (int k = 0; k < n_strikes; ++k) {
        const double *v = &acc[(size_t)k * TS_NV];
        b200mc_sums &o = out[k];
        o.n = (double)n_paths;
        o.sum_a = v[0]; o.sum_b = v[1]; o.sum_aa = v[2]; o.sum_bb = v[3]; o.sum_ab = v[4];
        o.sum_s = v[5]; o.sum_ss = v[6]; o.sum_ps = v[7];
    }
    return 0;
}

// Terminal values of paths driven by Sobol points with a caller-supplied bridge table: the device-side form of the
// reference's OWN use_sobol=True front end (engine/monte_carlo.py:290-299: generate_sobol_normals(n, 3 steps, seed),
// brownian_bridge_reorder on the first two blocks, the third block as jump sizes, jump uniforms from
// default_rng(seed + 1).random -- passed in as a host array).  With the reference's (degenerate) table this reproduces the
// reference's default price() inputs without its 7-11 s of host work per call.
extern "C" int b200mc_qmc_terminal(b200mc_handle *h, const b200mc_svj_params *p, double S0, double T, int32_t n_steps,
                                   int64_t n_paths, uint64_t path_offset, const uint32_t *sv, const uint32_t *shift,
                                   int32_t n_dims, int32_t bits, const b200mc_bridge_node *nodes, int32_t n_nodes,
                                   const double *Z_jump_host, const uint64_t *pcg64_state, uint32_t flags, double *S_final,
                                   double *S_anti)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    if (!p || !S_final) return fail(h, B200MC_EINVAL, "NULL argument");
    if (n_paths <= 0 || n_steps <= 0) return fail(h, B200MC_EINVAL, "n_paths and n_steps must be positive");
    if (!(T > 0.0) || !isfinite(T) || !isfinite(S0)) return fail(h, B200MC_EINVAL, "T must be positive and finite, S0 finite");
    if (path_offset + (uint64_t)n_paths > (1ull << (bits < 1 ? 1 : bits)))
        return fail(h, B200MC_EINVAL, "the Sobol sequence holds 2^bits points");
    const bool anti = flags & B200MC_ANTITHETIC;
    if (anti && !S_anti) return fail(h, B200MC_EINVAL, "B200MC_ANTITHETIC needs S_anti");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    const bool jumps = p->lambda_j * (T / n_steps) > 0.0;
    const bool own_u = Z_jump_host || pcg64_state;                                   // jump uniforms not from the point set
    const int nb_sobol = jumps ? (own_u ? 3 : 4) : (p->xi != 0.0 ? 2 : 1);           // blocks taken from the point set

    int64_t chunk = std::max<int64_t>(QMC_PT_MAX, ((int64_t)1 << 27) / n_steps);
    chunk = std::min(chunk, n_paths);
    const size_t zbytes = (size_t)chunk * n_steps * 8;
    const size_t extra_bytes = zbytes * 4 + (size_t)chunk * 8 * 3;                   // [Z1][Z2][Zjs][U][S][A][v]
    QmcPlan pl;
    char *extra = nullptr;
    B200MC_TRY(qmc_prepare(h, n_steps, sv, shift, n_dims, bits, nb_sobol, pl, extra_bytes, &extra, nodes, n_nodes));
    double *Z[4];
    for (int b = 0; b < 4; ++b) Z[b] = (double *)(extra + zbytes * b);
    double *dS = (double *)(extra + zbytes * 4), *dA = dS + chunk, *dV = dA + chunk;
    for (int64_t done = 0; done < n_paths; done += chunk) {
        const int64_t n = std::min(chunk, n_paths - done);
        if (jumps && Z_jump_host)
            B200MC_CUDA(h, cudaMemcpyAsync(Z[3], Z_jump_host + (size_t)done * n_steps, (size_t)n * n_steps * 8,
                                           cudaMemcpyHostToDevice, h->stream));
        else if (jumps && pcg64_state)       // row i of .random((n, steps)) = outputs [i steps, (i + 1) steps): one thread per row
            B200MC_TRY(pcg64_uniform_async(h, pcg64_state, (uint64_t)(path_offset + done) * (uint64_t)n_steps,
                                           n * (int64_t)n_steps, n_steps, Z[3]));
        for (int pass = 0; pass < (anti ? 2 : 1); ++pass) {
            const double sign = pass ? -1.0 : 1.0;                                   // monte_carlo.py:318-324
            for (int b = 0; b < nb_sobol; ++b)
                if (!(pass && b == 3)) B200MC_TRY(qmc_block(h, pl, b, path_offset + (uint64_t)done, n, sign, Z[b]));
            B200MC_TRY(b200mc_simulate_given_normals_dev(h, p, S0, T, n, n_steps, Z[0], nb_sobol >= 2 ? Z[1] : Z[0],
                                                         jumps ? Z[3] : Z[0], jumps ? Z[2] : Z[0], 0, pass ? dA : dS, dV,
                                                         nullptr));
        }
        B200MC_CUDA(h, cudaMemcpyAsync(S_final + done, dS, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
        if (anti) B200MC_CUDA(h, cudaMemcpyAsync(S_anti + done, dA, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return 0;
}
