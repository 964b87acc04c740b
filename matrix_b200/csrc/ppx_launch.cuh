// ppx_launch.cuh -- one kernel skeleton and one host launcher shared by every batch_* entry point.
//
// An "Op" is a trivially copyable functor holding the call's pointers and per-call constants; it is
// passed to the kernel as a __grid_constant__ parameter, i.e. it lives in the constant bank.  The
// kernel walks the batch in tiles of Op::BLOCK consecutive instances, one instance per thread, with
// a grid sized to a whole number of resident waves (148 SMs x blocks/SM on B200).
#pragma once

#include <atomic>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/ppx_cuda.h"
#include "ppx_io.cuh"
#include "ppx_math.cuh"

extern std::atomic<uint64_t> g_ppx_launch_count; // util.cu
extern std::atomic<int> g_ppx_square_lu;         // util.cu: ppx_cuda_set_square_solve_lu
int ppx_scratch_pool(int device, cudaMemPool_t *pool); // util.cu: the library's retained scratch pool of a device
int ppx_scratch_trim();                                // util.cu: give the retained scratch back to the driver

#define PPX_CHECK_PTRS(...)                          \
    do                                               \
    {                                                \
        const void *ptrs_[] = {__VA_ARGS__};         \
        for (const void *p_ : ptrs_)                 \
            if (p_ == nullptr)                       \
                return PPX_ERR_BAD_ARGUMENT;         \
    } while (0)


namespace ppx_dev
{

// An op may ask for a one-off start delay of the upper half of the CTA's warps (DESYNC_NS): kernels that alternate
// between phases bound by different pipes then run the two warps that share a scheduler in opposite phases.
template <class Op, class = void>
struct desync_ns : std::integral_constant<int, 0>
{
};
template <class Op>
struct desync_ns<Op, typename std::enable_if<(Op::DESYNC_NS > 0)>::type> : std::integral_constant<int, Op::DESYNC_NS>
{
};

template <int LAYOUT, class Op>
__global__ void __launch_bounds__(Op::BLOCK, Op::MINB) batch_kernel(const __grid_constant__ Op op, size_t n, size_t ld)
{
    extern __shared__ __align__(16) double ppx_smem[];
    if (desync_ns<Op>::value > 0 && (threadIdx.x >> 5) >= Op::BLOCK / 64)
        __nanosleep(desync_ns<Op>::value);
    const size_t stride = (size_t)gridDim.x * Op::BLOCK;
    for (size_t base = (size_t)blockIdx.x * Op::BLOCK; base < n; base += stride)
    {
        TileIO<LAYOUT> io(base, n, ld, ppx_smem, Op::MAXW);
        op(io);
    }
}

// Software-pipelined variant for ops that declare PREFETCH: the inputs of tile k+1 are requested with cp.async into
// a double-buffered shared-memory stage before tile k is computed, so the DRAM read latency of a tile is hidden
// behind the arithmetic of the previous one without holding any registers (a register prefetch made the
// 128-register kinematics kernel spill).  ncu on the unstaged kernel: 40 % of all warp samples sat in
// `long_scoreboard` on the first use of the freshly loaded joint angles.
// Op interface: IN_WIDTH (doubles per instance), prefetch(io, stage), compute_store(io, stage).
template <int LAYOUT, class Op>
__global__ void __launch_bounds__(Op::BLOCK, Op::MINB) batch_kernel_pf(const __grid_constant__ Op op, size_t n, size_t ld)
{
    extern __shared__ __align__(16) double ppx_smem[];
    constexpr size_t STAGE = TileIO<LAYOUT>::stage_doubles(Op::BLOCK, Op::IN_WIDTH);
    double *stage = ppx_smem + TileIO<LAYOUT>::smem_bytes(Op::BLOCK, Op::MAXW) / sizeof(double);
    const size_t stride = (size_t)gridDim.x * Op::BLOCK;
    size_t base = (size_t)blockIdx.x * Op::BLOCK;
    if (base >= n)
        return;
    int buf = 0;
    {
        TileIO<LAYOUT> io(base, n, ld, ppx_smem, Op::MAXW);
        op.prefetch(io, stage);
        cp_async_commit();
    }
    for (;;)
    {
        const size_t next = base + stride;
        const bool more = next < n;
        if (more)
        {
            TileIO<LAYOUT> ion(next, n, ld, ppx_smem, Op::MAXW);
            op.prefetch(ion, stage + (buf ^ 1) * STAGE);
        }
        cp_async_commit();
        cp_async_wait<1>(); // everything but the group just committed has landed: this tile's inputs
        TileIO<LAYOUT> io(base, n, ld, ppx_smem, Op::MAXW);
        op.compute_store(io, stage + buf * STAGE);
        if (!more)
            break;
        buf ^= 1;
        base = next;
    }
}

template <class Op, class = void>
struct has_prefetch : std::false_type
{
};
template <class Op>
struct has_prefetch<Op, typename std::enable_if<Op::PREFETCH>::type> : std::true_type
{
};

// an op may ask for the staged kernel on the record (AoS) layout only: PREFETCH_AOS_ONLY = true
template <class Op, class = void>
struct prefetch_aos_only : std::false_type
{
};
template <class Op>
struct prefetch_aos_only<Op, typename std::enable_if<Op::PREFETCH_AOS_ONLY>::type> : std::true_type
{
};
template <int LAYOUT, class Op>
struct use_prefetch
    : std::integral_constant<bool, has_prefetch<Op>::value && !(prefetch_aos_only<Op>::value && LAYOUT == LAYOUT_SOA)>
{
};

template <int LAYOUT, class Op>
constexpr typename std::enable_if<use_prefetch<LAYOUT, Op>::value, size_t>::type stage_bytes()
{
    return 2 * TileIO<LAYOUT>::stage_doubles(Op::BLOCK, Op::IN_WIDTH) * sizeof(double);
}
template <int LAYOUT, class Op>
constexpr typename std::enable_if<!use_prefetch<LAYOUT, Op>::value, size_t>::type stage_bytes()
{
    return 0;
}

template <int LAYOUT, class Op>
typename std::enable_if<use_prefetch<LAYOUT, Op>::value, void (*)(const Op, size_t, size_t)>::type pick_kernel()
{
    return batch_kernel_pf<LAYOUT, Op>;
}
template <int LAYOUT, class Op>
typename std::enable_if<!use_prefetch<LAYOUT, Op>::value, void (*)(const Op, size_t, size_t)>::type pick_kernel()
{
    return batch_kernel<LAYOUT, Op>;
}

struct DeviceInfo
{
    int sms = 0;
};

inline int device_info(DeviceInfo &info, int &device)
{
    static std::atomic<int> cache[64]; // SM count per device, 0 = not asked yet (host threads may race here)
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return (int)e;
    if (device < 0 || device >= 64)
        return PPX_ERR_NO_DEVICE;
    int sms = cache[device].load(std::memory_order_relaxed);
    if (sms == 0)
    {
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess)
            return (int)e;
        cache[device].store(sms, std::memory_order_relaxed);
    }
    info.sms = sms;
    return PPX_OK;
}

// Waves of resident CTAs the grid-stride loop is spread over: enough to hide the tail, few enough
// that block scheduling stays negligible.
constexpr int GRID_WAVES = 8;

template <int LAYOUT, class Op>
int launch(const Op &op, size_t n, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    if (LAYOUT == LAYOUT_SOA && ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    DeviceInfo info;
    int device = 0;
    int rc = device_info(info, device);
    if (rc != PPX_OK)
        return rc;
    auto kern = pick_kernel<LAYOUT, Op>();
    const size_t smem = TileIO<LAYOUT>::smem_bytes(Op::BLOCK, Op::MAXW) + stage_bytes<LAYOUT, Op>();
    static std::atomic<int> blocks_per_sm[64]; // per instantiation, per device; 0 = not asked yet.  Two host threads
                                               // may both ask (same answer, idempotent attribute), never tear
    int per_sm = blocks_per_sm[device].load(std::memory_order_relaxed);
    if (per_sm == 0)
    {
        if (smem > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return (int)e;
        }
        int b = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, Op::BLOCK, smem);
        if (e != cudaSuccess)
            return (int)e;
        per_sm = b > 0 ? b : 1;
        blocks_per_sm[device].store(per_sm, std::memory_order_relaxed);
    }
    const size_t tiles = (n + Op::BLOCK - 1) / Op::BLOCK;
    const size_t cap = (size_t)info.sms * per_sm * GRID_WAVES;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    kern<<<grid, Op::BLOCK, smem, (cudaStream_t)stream>>>(op, n, ld);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}

template <class Op>
int dispatch(int layout, const Op &op, size_t n, size_t ld, void *stream)
{
    if (layout == PPX_LAYOUT_AOS)
        return launch<LAYOUT_AOS>(op, n, ld, stream);
    if (layout == PPX_LAYOUT_SOA)
        return launch<LAYOUT_SOA>(op, n, ld, stream);
    return PPX_ERR_BAD_LAYOUT;
}

} // namespace ppx_dev
