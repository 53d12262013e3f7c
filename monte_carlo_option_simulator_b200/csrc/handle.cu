// handle.cu -- lifetime, error reporting, device-memory helpers and CUDA-event timing of libb200mc.
#include <new>

#include "common.cuh"

namespace b200mc {
thread_local char g_create_err[512] = "";
}
using namespace b200mc;

extern "C" int b200mc_version(void) { return B200MC_VERSION; }

extern "C" int b200mc_create(int device, b200mc_handle **out)
{
    if (!out) return fail(nullptr, B200MC_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, B200MC_ENODEVICE, "no CUDA device is visible (%s); libb200mc has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(nullptr, B200MC_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    B200MC_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return fail(nullptr, B200MC_ENODEVICE, "device %s is not an sm_100 part; the kernels are built for sm_100a only",
                    prop.name);
    }
    B200MC_CUDA(nullptr, cudaSetDevice(device));
    b200mc_handle *h = new (std::nothrow) b200mc_handle();
    if (!h) return fail(nullptr, B200MC_ENOMEM, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->cc = prop.major * 10 + prop.minor;
    h->hbm_bytes = prop.totalGlobalMem;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    h->sm_clock_khz = khz;
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaError_t e1 = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    cudaError_t e2 = cudaEventCreate(&h->ev0);
    cudaError_t e3 = cudaEventCreate(&h->ev1);
    cudaError_t e4 = cudaMalloc(&h->d_counter, 64);
    cudaError_t e5 = e4 == cudaSuccess ? cudaMemset(h->d_counter, 0, 64) : e4;
    h->own_stream = true;
    for (cudaError_t ee : {e1, e2, e3, e4, e5}) {
        if (ee != cudaSuccess) {
            fail(nullptr, B200MC_ECUDA, "handle set-up failed: %s", cudaGetErrorString(ee));
            b200mc_destroy(h);
            return B200MC_ECUDA;
        }
    }
    *out = h;
    return 0;
}

extern "C" int b200mc_destroy(b200mc_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    b200mc_peer_close(h);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_result) cudaFree(h->d_result);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_result) cudaFreeHost(h->h_result);
    if (h->h_pool) cudaFreeHost(h->h_pool);
    if (h->d_counter) cudaFree(h->d_counter);
    if (h->risk_state) cudaFree(h->risk_state);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" const char *b200mc_last_error(const b200mc_handle *h) { return h ? h->err : g_create_err; }

extern "C" int b200mc_device_info(const b200mc_handle *h, int *sm_count, int *sm_clock_khz, uint64_t *hbm_bytes,
                                  int *cc)
{
    if (!h) return B200MC_EINVAL;
    if (sm_count) *sm_count = h->sm_count;
    if (sm_clock_khz) *sm_clock_khz = h->sm_clock_khz;
    if (hbm_bytes) *hbm_bytes = h->hbm_bytes;
    if (cc) *cc = h->cc;
    return 0;
}

extern "C" int64_t b200mc_launch_count(const b200mc_handle *h) { return h ? h->launches : -1; }
extern "C" uint64_t b200mc_stream(const b200mc_handle *h) { return h ? (uint64_t)(uintptr_t)h->stream : 0; }

extern "C" int b200mc_set_stream(b200mc_handle *h, uint64_t stream)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    // only the handle's OWN stream is drained here: an external stream set earlier belongs to the caller (it may have
    // been destroyed already) and the caller orders its work itself
    if (h->own_stream) {
        B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaStreamDestroy(h->stream);
    }
    h->stream = (cudaStream_t)(uintptr_t)stream;
    h->own_stream = false;
    return 0;
}

extern "C" int b200mc_synchronize(b200mc_handle *h)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return peer_check(h);
}

extern "C" int b200mc_malloc(b200mc_handle *h, size_t bytes, void **dev_ptr)
{
    if (!h || !dev_ptr) return fail(h, B200MC_EINVAL, "NULL argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaMalloc(dev_ptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int b200mc_free(b200mc_handle *h, void *dev_ptr)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    B200MC_CUDA(h, cudaFree(dev_ptr));
    return 0;
}
extern "C" int b200mc_memcpy_h2d(b200mc_handle *h, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int b200mc_memcpy_d2h(b200mc_handle *h, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    B200MC_CUDA(h, cudaStreamSynchronize(h->stream));
    return peer_check(h);
}
extern "C" int b200mc_malloc_host(b200mc_handle *h, size_t bytes, void **host_ptr)
{
    if (!h || !host_ptr) return fail(h, B200MC_EINVAL, "NULL argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaMallocHost(host_ptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int b200mc_free_host(b200mc_handle *h, void *host_ptr)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaFreeHost(host_ptr));
    return 0;
}

extern "C" int b200mc_timer_begin(b200mc_handle *h)
{
    if (!h) return fail(nullptr, B200MC_EINVAL, "handle is NULL");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    return 0;
}
extern "C" int b200mc_timer_end(b200mc_handle *h, float *elapsed_ms)
{
    if (!h || !elapsed_ms) return fail(h, B200MC_EINVAL, "NULL argument");
    B200MC_CUDA(h, cudaSetDevice(h->device));
    B200MC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    B200MC_CUDA(h, cudaEventSynchronize(h->ev1));
    B200MC_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return 0;
}
