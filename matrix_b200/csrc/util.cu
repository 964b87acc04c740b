// util.cu -- device/memory helpers behind the C ABI, layout conversion, the counter-based input
// generator and the fp64 peak probe.  None of this is on the timed hot path.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include <sys/mman.h>

#include "ppx_launch.cuh"

std::atomic<uint64_t> g_ppx_launch_count{0};
std::atomic<int> g_ppx_square_lu{1};

// One memory pool per device for the library's own short-lived device scratch (host-ABI staging buffers, the work
// counters of the persistent kernels).  Unlike the default pool it keeps up to PPX_HOST_POOL_MIB (default 1024) MiB of
// freed memory mapped, so that a call does not pay for mapping device memory again.
namespace
{
std::mutex g_pool_mu;
cudaMemPool_t g_pools[64] = {};
} // namespace

int ppx_scratch_pool(int device, cudaMemPool_t *pool)
{
    if (device < 0 || device >= 64)
        return PPX_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (g_pools[device] == nullptr)
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&g_pools[device], &props);
        if (e != cudaSuccess)
        {
            g_pools[device] = nullptr;
            return (int)e;
        }
        const char *env = std::getenv("PPX_HOST_POOL_MIB");
        const long mib = env ? std::atol(env) : 1024;
        unsigned long long keep = (unsigned long long)(mib < 0 ? 0 : mib) << 20;
        cudaMemPoolSetAttribute(g_pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *pool = g_pools[device];
    return PPX_OK;
}
int ppx_scratch_trim()
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (cudaMemPool_t p : g_pools)
        if (p != nullptr)
        {
            cudaError_t e = cudaMemPoolTrimTo(p, 0);
            if (e != cudaSuccess)
                return (int)e;
        }
    return PPX_OK;
}

namespace
{

// records of `width` doubles, dense AoS <-> component-major SoA; one thread per (instance, component),
// coalesced on the SoA side, L1/L2-merged on the AoS side (a conversion utility, not a hot kernel)
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, int width, size_t n, size_t ld)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        for (int k = 0; k < width; k++)
            soa[(size_t)k * ld + i] = aos[i * width + k];
}

__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, int width, size_t n, size_t ld)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        for (int k = 0; k < width; k++)
            aos[i * width + k] = soa[(size_t)k * ld + i];
}

// splitmix64 finaliser over the counter seed + (j+1)*GOLDEN: matrix_b200/synth.py, bit for bit
__global__ void k_fill_uniform(double *__restrict__ out, size_t count, uint64_t seed, uint64_t first, double lo,
                               double hi)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double span = hi - lo;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride)
    {
        uint64_t z = seed + (first + j + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        out[j] = __dadd_rn(lo, __dmul_rn(span, u)); // no FMA: numpy computes lo + span*u in two roundings
    }
}

// 8 independent DFMA chains per thread, enough resident warps to fill every fp64 pipe
__global__ void __launch_bounds__(256) k_fp64_probe(double *sink, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++)
    {
        a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
        a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456)
        sink[0] = s;
}

// measurement helper: the store side of the SoA kernels in isolation -- `planes` component planes of n doubles,
// thread i writing element i of every plane (mode 0: one 64-bit store per plane, like the batch kernels; mode 1: two
// instances per thread, one 128-bit store per plane)
template <int MODE>
__global__ void k_probe_write(double *__restrict__ out, int planes, size_t n, size_t ld)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (MODE == 0)
    {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            for (int k = 0; k < planes; k++)
                __stcs(out + (size_t)k * ld + i, (double)k);
    }
    else
    {
        for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n; i += 2 * stride)
            for (int k = 0; k < planes; k++)
                __stcs(reinterpret_cast<double2 *>(out + (size_t)k * ld + i), make_double2((double)k, (double)k));
    }
}

// blocks handed out by ppx_cuda_malloc_host's mapped path: user pointer -> (mapping, its length, registered length)
struct HostBlock
{
    void *raw;
    size_t raw_len;
    size_t len;
};
std::mutex g_hb_mu;
std::unordered_map<void *, HostBlock> g_hb;

int grid_for(size_t n, int block)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t tiles = (n + block - 1) / block;
    size_t cap = (size_t)sms * 16;
    return (int)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

} // namespace

extern "C"
{

const char *ppx_cuda_version(void) { return "ppx-b200 0.2 (sm_100a, fp64, reference ppx v1.2)"; }

const char *ppx_cuda_error_string(int code)
{
    switch (code)
    {
    case PPX_OK: return "ok";
    case PPX_ERR_BAD_LAYOUT: return "layout must be PPX_LAYOUT_AOS or PPX_LAYOUT_SOA";
    case PPX_ERR_BAD_SHAPE: return "shape not compiled into libppx_cuda";
    case PPX_ERR_BAD_ARGUMENT: return "bad argument (null pointer, ld < n, out-of-range parameter)";
    case PPX_ERR_NO_DEVICE: return "no usable CUDA device";
    case PPX_ERR_HOST: return "host-side resource failure (worker thread or host memory)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown ppx error";
    }
}

int ppx_cuda_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int ppx_cuda_set_device(int device) { return (int)cudaSetDevice(device); }
int ppx_cuda_malloc(void **dptr, size_t bytes) { return (int)cudaMalloc(dptr, bytes); }
int ppx_cuda_free(void *dptr) { return (int)cudaFree(dptr); }
// Pinned host memory.  Small blocks come from cudaHostAlloc.  Large ones (>= 64 MiB) are mapped here, advised to use
// transparent huge pages (PPX_HOST_HUGEPAGES=0 turns that off), faulted in by several threads at once -- a single
// thread zero-filling 4 KiB pages manages ~1-2 GB/s, which is what made an 8 GB cudaHostAlloc take seconds -- and then
// page-locked with ONE cudaHostRegister(Portable), so that a copy may span any part of the block and every device of
// the process can DMA to it.
int ppx_cuda_malloc_host(void **hptr, size_t bytes)
{
    if (!hptr)
        return PPX_ERR_BAD_ARGUMENT;
    *hptr = nullptr;
    if (bytes == 0)
        return PPX_OK;
    const char *mode = std::getenv("PPX_HOST_ALLOC"); // "cuda": always cudaHostAlloc
    const bool map_it = bytes >= ((size_t)64 << 20) && !(mode && std::strcmp(mode, "cuda") == 0);
    if (map_it)
    {
        constexpr size_t HUGE = (size_t)2 << 20;
        const size_t len = (bytes + HUGE - 1) / HUGE * HUGE;
        void *raw = mmap(nullptr, len + HUGE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (raw != MAP_FAILED)
        {
            char *p = (char *)(((uintptr_t)raw + HUGE - 1) / HUGE * HUGE);
            const char *hp = std::getenv("PPX_HOST_HUGEPAGES");
            if (!(hp && std::atoi(hp) == 0))
                madvise(p, len, MADV_HUGEPAGE);
            unsigned nt = std::thread::hardware_concurrency();
            nt = nt == 0 ? 4 : (nt > 16 ? 16 : nt);
            const size_t slices = std::min<size_t>(nt, len / ((size_t)32 << 20) + 1);
            std::vector<std::thread> th;
            try // the parallel first touch is an optimisation: pages it does not reach are faulted by the registration
            {
                th.reserve(slices);
                for (size_t t = 0; t < slices; t++)
                    th.emplace_back([=] {
                        const size_t lo = len / slices * t, hi = t + 1 == slices ? len : len / slices * (t + 1);
                        for (size_t o = lo; o < hi; o += 4096)
                            ((volatile char *)p)[o] = 0;
                    });
            }
            catch (...)
            {
            }
            for (std::thread &t : th)
                t.join();
            const cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
            if (e == cudaSuccess)
            {
                try
                {
                    std::lock_guard<std::mutex> lock(g_hb_mu);
                    g_hb[p] = HostBlock{raw, len + HUGE, len};
                }
                catch (...)
                {
                    cudaHostUnregister(p);
                    munmap(raw, len + HUGE);
                    return PPX_ERR_HOST;
                }
                *hptr = p;
                return PPX_OK;
            }
            (void)cudaGetLastError(); // registration refused (locked-memory limit, ...): fall back to the runtime's allocator
            munmap(raw, len + HUGE);
        }
    }
    return (int)cudaHostAlloc(hptr, bytes, cudaHostAllocPortable);
}
int ppx_cuda_free_host(void *hptr)
{
    if (!hptr)
        return PPX_OK;
    HostBlock b{};
    bool mapped = false;
    {
        std::lock_guard<std::mutex> lock(g_hb_mu);
        auto it = g_hb.find(hptr);
        if (it != g_hb.end())
        {
            b = it->second;
            mapped = true;
            g_hb.erase(it);
        }
    }
    if (!mapped)
        return (int)cudaFreeHost(hptr);
    const cudaError_t e = cudaHostUnregister(hptr);
    munmap(b.raw, b.raw_len);
    return (int)e;
}

int ppx_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}

int ppx_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}

int ppx_cuda_stream_synchronize(void *stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }

uint64_t ppx_cuda_launch_count(void) { return g_ppx_launch_count.load(std::memory_order_relaxed); }
int ppx_cuda_set_square_solve_lu(int enable) { return g_ppx_square_lu.exchange(enable != 0 ? 1 : 0); }

int ppx_cuda_aos_to_soa(const double *aos, double *soa, int width, size_t n, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    if (!aos || !soa || width < 1 || ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    k_aos_to_soa<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(aos, soa, width, n, ld);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}

int ppx_cuda_soa_to_aos(const double *soa, double *aos, int width, size_t n, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    if (!aos || !soa || width < 1 || ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    k_soa_to_aos<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(soa, aos, width, n, ld);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}

int ppx_cuda_fill_uniform(double *out, size_t count, uint64_t seed, uint64_t first, double lo, double hi,
                          void *stream)
{
    if (count == 0)
        return PPX_OK;
    if (!out)
        return PPX_ERR_BAD_ARGUMENT;
    k_fill_uniform<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(out, count, seed, first, lo, hi);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}

int ppx_cuda_probe_write(double *out, int planes, size_t n, size_t ld, int mode, void *stream)
{
    if (!out || planes < 1 || ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    if (mode == 0)
        k_probe_write<0><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(out, planes, n, ld);
    else
        k_probe_write<1><<<grid_for(n / 2, 256), 256, 0, (cudaStream_t)stream>>>(out, planes, n, ld);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}

int ppx_cuda_fp64_peak_probe(double *tflops, void *stream)
{
    if (!tflops)
        return PPX_ERR_BAD_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *sink = nullptr;
    if ((e = cudaMalloc(&sink, sizeof(double))) != cudaSuccess)
        return (int)e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    k_fp64_probe<<<blocks, threads, 0, s>>>(sink, 1 << 10); // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++)
    {
        cudaEventRecord(t0, s);
        k_fp64_probe<<<blocks, threads, 0, s>>>(sink, iters);
        cudaEventRecord(t1, s);
        cudaEventSynchronize(t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
        const double tf = flops / ((double)ms * 1e-3) * 1e-12;
        if (tf > best)
            best = tf;
    }
    g_ppx_launch_count.fetch_add(6, std::memory_order_relaxed);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(sink);
    *tflops = best;
    return (int)cudaGetLastError();
}

} // extern "C"
