// matmul.cu -- batched MatrixS<M,N>::operator* (include/matrixs.hpp:596-640 of the reference): C = A B for
// millions of independent fixed-size pairs, one instance per thread.
//
// The reference zero-initialises C and accumulates in k-j-i order, c(i,j) += a(i,k) b(k,j); every element of C
// therefore receives its N products in increasing k, each product rounded, each addition rounded.  The kernels do
// exactly that with explicit __dmul_rn / __dadd_rn (no contraction into FMA), so the result is BIT-IDENTICAL to the
// reference built without FMA contraction (oracle/_ref/libppx_ref_strict.so); against the reference's default x86
// build, where GCC contracts a*b+c, it differs by the usual one rounding per term.  The op is HBM-bound at every
// shape on the menu (6x6x6: 432 fp64 instructions on 864 B), so leaving the FMA out costs nothing.
//
// Shapes on the menu are fully unrolled in registers; any other shape with all dimensions <= 8 runs the generic
// kernel (runtime loops, operands re-read through L1), so ppx::Matrix<M,N>::operator* never needs a CPU loop.
#include "ppx_launch.cuh"

using namespace ppx_dev;

namespace
{

template <int M, int N, int L>
struct OpMatmul
{
    static constexpr int W1 = M * N > N * L ? M * N : N * L;
    static constexpr int BLOCK = 128, MINB = (M * N + N * L + M * L <= 48 ? 4 : 2), MAXW = W1 > M * L ? W1 : M * L;
    const double *A;
    const double *B;
    double *C;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], b[N * L], c[M * L];
        io.load(A, a);
        io.load(B, b);
#pragma unroll
        for (int i = 0; i < M * L; i++)
            c[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; k++)
#pragma unroll
            for (int j = 0; j < L; j++)
#pragma unroll
                for (int i = 0; i < M; i++)
                    c[i + j * M] = __dadd_rn(c[i + j * M], __dmul_rn(a[i + k * M], b[k + j * N]));
        io.store(C, c);
    }
};

// any (m, n, l) with all dimensions <= 8: one instance per thread, operands read straight from global memory (the
// re-reads of A hit L1), same summation order per element as above
template <int LAYOUT>
__global__ void __launch_bounds__(128) matmul_dyn_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                         double *__restrict__ C, int m, int n, int l, size_t count,
                                                         size_t ld)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        for (int j = 0; j < l; j++)
            for (int r = 0; r < m; r++)
            {
                double acc = 0.0;
                for (int k = 0; k < n; k++)
                {
                    const double a = LAYOUT == LAYOUT_AOS ? A[i * (size_t)(m * n) + r + k * m] : A[(size_t)(r + k * m) * ld + i];
                    const double b = LAYOUT == LAYOUT_AOS ? B[i * (size_t)(n * l) + k + j * n] : B[(size_t)(k + j * n) * ld + i];
                    acc = __dadd_rn(acc, __dmul_rn(a, b));
                }
                if (LAYOUT == LAYOUT_AOS)
                    C[i * (size_t)(m * l) + r + j * m] = acc;
                else
                    C[(size_t)(r + j * m) * ld + i] = acc;
            }
    }
}

} // namespace

#define PPX_SHAPE3(M_, N_, L_)               \
    if (m == M_ && n == N_ && l == L_)       \
    return dispatch(layout, OpMatmul<M_, N_, L_>{A, B, C}, count, ld, stream)

extern "C" int ppx_cuda_batch_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count,
                                     int layout, size_t ld, void *stream)
{
    if (m < 1 || n < 1 || l < 1 || m > 8 || n > 8 || l > 8)
        return PPX_ERR_BAD_SHAPE;
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, B, C);
    PPX_SHAPE3(3, 3, 3);
    PPX_SHAPE3(4, 4, 4);
    PPX_SHAPE3(6, 6, 6);
    PPX_SHAPE3(6, 6, 1);
    PPX_SHAPE3(4, 4, 1);
    PPX_SHAPE3(3, 3, 1);
    if (layout != PPX_LAYOUT_AOS && layout != PPX_LAYOUT_SOA)
        return PPX_ERR_BAD_LAYOUT;
    if (layout == PPX_LAYOUT_SOA && ld < count)
        return PPX_ERR_BAD_ARGUMENT;
    DeviceInfo info;
    int device = 0;
    const int rc = device_info(info, device);
    if (rc != PPX_OK)
        return rc;
    const size_t tiles = (count + 127) / 128, cap = (size_t)info.sms * 16;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    if (layout == PPX_LAYOUT_AOS)
        matmul_dyn_kernel<LAYOUT_AOS><<<grid, 128, 0, (cudaStream_t)stream>>>(A, B, C, m, n, l, count, ld);
    else
        matmul_dyn_kernel<LAYOUT_SOA><<<grid, 128, 0, (cudaStream_t)stream>>>(A, B, C, m, n, l, count, ld);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaPeekAtLastError();
}
