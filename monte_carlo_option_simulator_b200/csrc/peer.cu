// peer.cu -- the one exchange step of the multi-GPU path (the all-reduce of the per-strike sum vectors,
// 136 bytes per strike) as a one-shot all-reduce over NVLink peer memory instead of an NCCL call.
//
// Every rank owns an exchange buffer in its HBM, exported with CUDA IPC and mapped by all peers (one process per GPU).
// k_peer_allreduce, one CTA, enqueued on the handle's stream right behind the kernel that produced the sums:
//   1. stores this rank's vector into slot [parity][rank] of EVERY rank's buffer (peer stores travel over NVLink /
//      NVSwitch), fences system-wide, then raises flag [parity][rank] = epoch on every rank;
//   2. spins until the flags of all ranks for this epoch have arrived in its own buffer;
//   3. adds the slots in rank order (so every rank gets bitwise the same sums) over the caller's vector.
// Two parities: a rank can be at most one collective ahead of a peer (it needs the peer's flag to finish its own), so
// slot [parity] of epoch e is not overwritten before every reader of epoch e - 2 is done.  The message is a few hundred
// bytes, so the cost is one NVLink store round (~2-3 us) instead of a collective launch plus its protocol (~10-25 us
// measured for ncclAllReduce of 136 B inside bench.py's timed region).  All ranks must call in the same order.
#include "common.cuh"

namespace b200mc {

struct PeerBuf {
    unsigned long long flag[2][B200MC_PEER_MAX_RANKS];
    double data[2][B200MC_PEER_MAX_RANKS][B200MC_PEER_MAX_DOUBLES];
};

struct PeerArgs {
    PeerBuf *peer[B200MC_PEER_MAX_RANKS];
    int rank, world;
    unsigned long long epoch;
    unsigned long long *status;      // device alias of the handle's pinned status word: epoch of the first timed-out exchange
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// T = double (sums) or unsigned long long (histogram counts of the sharded radix select, risk.cu): 8-byte elements
template <typename T>
__global__ void __launch_bounds__(256)
k_peer_allreduce(const __grid_constant__ PeerArgs a, T *__restrict__ data, int n)
{
    const int tid = threadIdx.x, par = (int)(a.epoch & 1ull);
    __shared__ int timed_out;
    if (tid == 0) timed_out = 0;
    for (int r = 0; r < a.world; ++r) {
        T *dst = reinterpret_cast<T *>(a.peer[r]->data[par][a.rank]);
        for (int i = tid; i < n; i += blockDim.x) dst[i] = data[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) st_release_sys(&a.peer[tid]->flag[par][a.rank], a.epoch);
    if (tid < a.world) {
        const unsigned long long *f = &a.peer[a.rank]->flag[par][tid];
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) < a.epoch) {
            if (global_ns() - t0 > 20000000000ull) {      // 20 s: a peer never arrived -- report, never hang the GPU
                timed_out = 1;
                if (*(volatile unsigned long long *)a.status == 0ull) {     // pinned host memory, read by the host at
                    *(volatile unsigned long long *)a.status = a.epoch;     // its next synchronisation point
                    __threadfence_system();
                }
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const PeerBuf *mine = a.peer[a.rank];
    for (int i = tid; i < n; i += blockDim.x) {
        T s = (T)0;
        for (int r = 0; r < a.world; ++r) s += __ldcg(reinterpret_cast<const T *>(&mine->data[par][r][i]));   // written by peers: read at L2
        // a timed-out exchange never yields a plausible number: NaN for the sums, all ones for the histogram counts; the
        // host turns the status word into B200MC_ECUDA before any of it is used (peer_check)
        if (timed_out) {
            if constexpr (sizeof(T) == 8 && T(0.5) != T(0)) s = (T)__longlong_as_double(0x7ff8000000000000ll);
            else s = (T)~0ull;
        }
        data[i] = s;
    }
}

// Host side of the time-out report: called after every stream synchronisation that follows an exchange.
int peer_check(b200mc_handle *h)
{
    if (!h->peer_status) return 0;
    const unsigned long long e = *(volatile unsigned long long *)h->peer_status;
    if (e == 0ull) return 0;
    char num[32];
    snprintf(num, sizeof(num), "%llu", e);
    return fail(h, B200MC_ECUDA, "peer all-reduce %s timed out after 20 s: a rank never entered the collective "
                                 "(results of this and later exchanges are invalid; reconnect the peers)", num);
}

int peer_allreduce_async(b200mc_handle *h, void *data_dev, int32_t n, bool as_u64)
{
    if (h->peer_world < 1) return fail(h, B200MC_EINVAL, "call b200mc_peer_connect first");
    if (n < 1 || n > B200MC_PEER_MAX_DOUBLES)
        return fail(h, B200MC_EINVAL, "n_doubles must be in [1, 4352] (256 strikes x 17 sums)");
    PeerArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < h->peer_world; ++r) a.peer[r] = (PeerBuf *)h->peer_ptr[r];
    a.rank = h->peer_rank;
    a.world = h->peer_world;
    a.epoch = ++h->peer_epoch;
    a.status = (unsigned long long *)h->peer_status_dev;
    if (as_u64) k_peer_allreduce<unsigned long long><<<1, 256, 0, h->stream>>>(a, (unsigned long long *)data_dev, n);
    else k_peer_allreduce<double><<<1, 256, 0, h->stream>>>(a, (double *)data_dev, n);
    B200MC_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

} // namespace b200mc

using namespace b200mc;

extern "C" int b200mc_peer_create(b200mc_handle *h, unsigned char ipc_handle_out[64])
{
    if (!h || !ipc_handle_out) return fail(h, B200MC_EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    if (h->peer_world > 0) return fail(h, B200MC_EINVAL, "peers are connected: call b200mc_peer_close first");
    if (!h->peer_local) {
        B200MC_CUDA(h, cudaMalloc(&h->peer_local, sizeof(PeerBuf)));
    }
    // a (re)created exchange starts from zeroed flags and epoch 0 (stale flags of an earlier connection would satisfy
    // the waits of the new one immediately)
    B200MC_CUDA(h, cudaMemset(h->peer_local, 0, sizeof(PeerBuf)));
    B200MC_CUDA(h, cudaDeviceSynchronize());
    if (!h->peer_status) {
        B200MC_CUDA(h, cudaHostAlloc(&h->peer_status, 64, cudaHostAllocMapped));
        B200MC_CUDA(h, cudaHostGetDevicePointer(&h->peer_status_dev, h->peer_status, 0));
    }
    memset(h->peer_status, 0, 64);
    cudaIpcMemHandle_t ih;
    B200MC_CUDA(h, cudaIpcGetMemHandle(&ih, h->peer_local));
    memcpy(ipc_handle_out, &ih, 64);
    return 0;
}

extern "C" int b200mc_peer_connect(b200mc_handle *h, int rank, int world, const unsigned char *all_handles)
{
    if (!h || !all_handles) return fail(h, B200MC_EINVAL, "NULL argument");
    if (!h->peer_local) return fail(h, B200MC_EINVAL, "call b200mc_peer_create first");
    if (world < 1 || world > B200MC_PEER_MAX_RANKS || rank < 0 || rank >= world)
        return fail(h, B200MC_EINVAL, "rank / world out of range (at most 16 ranks)");
    if (h->peer_world > 0) return fail(h, B200MC_EINVAL, "already connected: call b200mc_peer_close, then b200mc_peer_create again");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->peer_ptr[r] = h->peer_local; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, all_handles + (size_t)r * 64, 64);
        void *p = nullptr;
        B200MC_CUDA(h, cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        h->peer_ptr[r] = p;
    }
    h->peer_rank = rank;
    h->peer_world = world;
    h->peer_epoch = 0;
    return 0;
}

extern "C" int b200mc_peer_allreduce(b200mc_handle *h, double *data_dev, int32_t n_doubles)
{
    if (!h || !data_dev) return fail(h, B200MC_EINVAL, "NULL argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    return peer_allreduce_async(h, data_dev, n_doubles, false);
}

extern "C" int b200mc_peer_close(b200mc_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < h->peer_world; ++r)
        if (r != h->peer_rank && h->peer_ptr[r]) cudaIpcCloseMemHandle(h->peer_ptr[r]);
    if (h->peer_local) cudaFree(h->peer_local);
    if (h->peer_status) cudaFreeHost(h->peer_status);
    h->peer_local = nullptr;
    h->peer_status = nullptr;
    h->peer_status_dev = nullptr;
    h->peer_world = 0;
    h->peer_epoch = 0;
    memset(h->peer_ptr, 0, sizeof(h->peer_ptr));
    return 0;
}
