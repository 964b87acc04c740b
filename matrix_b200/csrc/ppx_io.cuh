// ppx_io.cuh -- how a batch of fixed-size records moves between HBM and the registers of the
// thread that owns the instance.
//
//   SOA  component-major, base[k*ld + i].  Component k of 32 consecutive instances is 256
//        contiguous bytes, so every warp-level access is two full 128-byte lines: coalesced by
//        construction, no staging.  Loads use ld.global.cs / stores st.global.cs (evict-first):
//        every byte is touched exactly once per launch and must not displace anything in L2.
//   AOS  dense reference records, base[i*D + k] (std::vector<ppx::MatrixS<M,N>>::data()).  A
//        thread-per-record access would touch 32 different sectors per instruction, so each warp
//        moves its 32 records (one contiguous 32*D*8-byte span) with coalesced accesses through a
//        private shared-memory tile whose record stride is padded to an odd number of doubles;
//        the per-thread side of the transpose is then bank-conflict free.  Only __syncwarp is
//        needed: warps never wait for each other.
#pragma once

#include <cstddef>
#include <cstdint>

namespace ppx_dev
{

// store / load flavour, selectable at build time for experiments (-DPPX_STORE=0 plain, 1 .cs evict-first [default],
// 2 .wt write-through; -DPPX_LOAD likewise)
#ifndef PPX_STORE
#define PPX_STORE 1
#endif
__device__ __forceinline__ void stg(double *p, double v)
{
#if PPX_STORE == 0
    *p = v;
#elif PPX_STORE == 2
    __stwt(p, v);
#elif PPX_STORE == 3
    __stcg(p, v);
#else
    __stcs(p, v);
#endif
}

constexpr int LAYOUT_AOS = 0;
constexpr int LAYOUT_SOA = 1;

// cp.async (LDGSTS): global -> shared without passing through registers, 8 bytes per request
__device__ __forceinline__ void cp_async_8(double *smem_dst, const double *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

template <int LAYOUT>
struct TileIO;

template <>
struct TileIO<LAYOUT_SOA>
{
    size_t i;   // this thread's instance
    size_t ld;  // component stride
    bool valid; // i < n

    __host__ __device__ static constexpr size_t smem_bytes(int /*block*/, int /*max_width*/) { return 0; }

    __device__ __forceinline__ TileIO(size_t tile_base, size_t n, size_t ld_, double * /*smem*/, int /*max_width*/)
        : i(tile_base + threadIdx.x), ld(ld_), valid(tile_base + threadIdx.x < n) {}

    template <int D>
    __device__ __forceinline__ void load(const double *__restrict__ base, double (&r)[D]) const
    {
#pragma unroll
        for (int k = 0; k < D; k++)
            r[k] = valid ? __ldcs(base + (size_t)k * ld + i) : 0.0;
    }

    template <int D>
    __device__ __forceinline__ void store(double *__restrict__ base, const double (&r)[D]) const
    {
        if (valid)
        {
#pragma unroll
            for (int k = 0; k < D; k++)
                stg(base + (size_t)k * ld + i, r[k]);
        }
    }

    // piecewise store of a D-wide record: components off .. off+CNT-1 now, flush<D>() once at the end
    template <int D, int CNT>
    __device__ __forceinline__ void put(double *__restrict__ base, int off, const double (&r)[CNT]) const
    {
        if (valid)
        {
#pragma unroll
            for (int k = 0; k < CNT; k++)
                stg(base + (size_t)(off + k) * ld + i, r[k]);
        }
    }

    template <int D>
    __device__ __forceinline__ void flush(double *__restrict__ /*base*/) const {}

    // asynchronous input staging: this thread's D components go to stage[k*blockDim + tid] and are read back by
    // the same thread after cp_async_wait (no block-level synchronisation needed)
    __host__ __device__ static constexpr size_t stage_doubles(int block, int width) { return (size_t)block * width; }
    template <int D>
    __device__ __forceinline__ void async_load(const double *__restrict__ base, double *stage) const
    {
        if (valid)
        {
#pragma unroll
            for (int k = 0; k < D; k++)
                cp_async_8(stage + k * blockDim.x + threadIdx.x, base + (size_t)k * ld + i);
        }
    }
    template <int D>
    __device__ __forceinline__ void read_staged(const double *stage, double (&r)[D]) const
    {
#pragma unroll
        for (int k = 0; k < D; k++)
            r[k] = valid ? stage[k * blockDim.x + threadIdx.x] : 0.0;
    }
};

template <>
struct TileIO<LAYOUT_AOS>
{
    size_t i;     // this thread's instance
    size_t i0;    // first instance of this warp's 32-record span
    int nvalid;   // records of the span that exist (0..32)
    int lane;
    bool valid;
    double *ws;   // this warp's staging tile: 32 records, stride (max_width | 1) doubles

    __host__ __device__ static constexpr size_t smem_bytes(int block, int max_width)
    {
        return (size_t)(block / 32) * 32 * (size_t)(max_width | 1) * sizeof(double);
    }

    __device__ __forceinline__ TileIO(size_t tile_base, size_t n, size_t /*ld*/, double *smem, int max_width)
    {
        lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        i0 = tile_base + (size_t)warp * 32;
        i = i0 + lane;
        valid = i < n;
        const size_t left = i0 < n ? n - i0 : 0;
        nvalid = left < 32 ? (int)left : 32;
        ws = smem + (size_t)warp * 32 * (max_width | 1);
    }

    template <int D>
    __device__ __forceinline__ void load(const double *__restrict__ base, double (&r)[D]) const
    {
        constexpr int DP = D | 1;
        const double *src = base + i0 * D;
        const int total = nvalid * D;
        // coalesced: lane l reads doubles l, l+32, ... of the warp's contiguous span
#pragma unroll
        for (int m = 0; m < D; m++)
        {
            const int j = lane + 32 * m;
            if (j < total)
            {
                const int rec = j / D;
                ws[rec * DP + (j - rec * D)] = __ldcs(src + j);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < D; k++)
            r[k] = valid ? ws[lane * DP + k] : 0.0;
        __syncwarp();
    }

    template <int D, int CNT>
    __device__ __forceinline__ void put(double *__restrict__ /*base*/, int off, const double (&r)[CNT]) const
    {
        constexpr int DP = D | 1;
#pragma unroll
        for (int k = 0; k < CNT; k++)
            ws[lane * DP + off + k] = r[k];
    }

    template <int D>
    __device__ __forceinline__ void store(double *__restrict__ base, const double (&r)[D]) const
    {
        put<D, D>(base, 0, r);
        flush<D>(base);
    }

    template <int D>
    __device__ __forceinline__ void flush(double *__restrict__ base) const
    {
        constexpr int DP = D | 1;
        __syncwarp();
        double *dst = base + i0 * D;
        const int total = nvalid * D;
#pragma unroll
        for (int m = 0; m < D; m++)
        {
            const int j = lane + 32 * m;
            if (j < total)
            {
                const int rec = j / D;
                stg(dst + j, ws[rec * DP + (j - rec * D)]);
            }
        }
        __syncwarp();
    }

    // asynchronous input staging: the warp's contiguous span goes, element by element, into a padded tile of the
    // stage buffer; read_staged() is the per-thread side after cp_async_wait + __syncwarp
    __host__ __device__ static constexpr size_t stage_doubles(int block, int width) { return (size_t)block * (width | 1); }
    template <int D>
    __device__ __forceinline__ void async_load(const double *__restrict__ base, double *stage) const
    {
        constexpr int DP = D | 1;
        double *tile = stage + (size_t)(threadIdx.x >> 5) * 32 * DP;
        const double *src = base + i0 * D;
        const int total = nvalid * D;
#pragma unroll
        for (int m = 0; m < D; m++)
        {
            const int j = lane + 32 * m;
            if (j < total)
            {
                const int rec = j / D;
                cp_async_8(tile + rec * DP + (j - rec * D), src + j);
            }
        }
    }
    template <int D>
    __device__ __forceinline__ void read_staged(const double *stage, double (&r)[D]) const
    {
        constexpr int DP = D | 1;
        const double *tile = stage + (size_t)(threadIdx.x >> 5) * 32 * DP;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < D; k++)
            r[k] = valid ? tile[lane * DP + k] : 0.0;
    }
};

} // namespace ppx_dev
