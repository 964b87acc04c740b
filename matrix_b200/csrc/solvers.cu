// solvers.cu -- batched least-squares QR, SVD / pseudo-inverse solve and symmetric eigensolver
// (include/linalg.hpp of the reference: QR<m,n> :535-603, SVD<m,n> :605-924, EigenValue<n> :967-1167,
// linsolve<> :1169-1209), one instance per thread, everything in registers.
//
// QR follows the reference's Householder recurrence step for step (fully unrolled).  SVD and eig do
// NOT port the reference's Golub-Kahan-Reinsch / tred2+QL iterations -- their data-dependent deflation
// logic and dynamically indexed work arrays would put a 6x6 problem in local memory and diverge inside
// every warp.  They use cyclic Jacobi rotations instead: fixed, fully unrolled rotation schedules on
// register-resident matrices, sweeps terminated by a warp vote.  The factors agree with the reference's
// up to the order/sign freedom of the decomposition; singular values, eigenvalues, reconstruction and
// every solve agree to fp64 rounding (parity tests: tests/test_gpu_parity.py).
#include "ppx_launch.cuh"

using namespace ppx_dev;

namespace
{

// BASELINE config 3 at (4,3): 128 B in, 24 B + 1 B out, ~250 fp64 instructions: HBM-bound.
template <int M, int N>
struct OpLinsolveQR
{
    static constexpr int BLOCK = 128, MINB = (M * N <= 16 ? 4 : 2), MAXW = M * N;
    const double *A;
    const double *b;
    double *x;
    int8_t *status;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], rhs[M], sol[N];
        io.load(A, a);
        io.load(b, rhs);
        const int st = linsolve_qr<M, N>(a, rhs, sol);
        io.store(x, sol);
        if (status != nullptr && io.valid)
            status[io.i] = (int8_t)st;
    }
};

// BASELINE config 4a at (6,6): 288 B in, 624 B out, ~9k fp64 instructions: fp64-bound.
template <int M, int N, bool WANT_U, bool WANT_V>
struct OpSVD
{
    // 6x6: one CTA of 8 warps per SM (255 registers per thread).  Measured with the preconditioned iteration:
    // 1.99 G inst/s against 1.79 G for two CTAs of 4 warps (same occupancy, same register cap).
#ifndef PPX_SVD_MINB
#define PPX_SVD_MINB 1
#endif
#ifndef PPX_SVD_BLOCK
#define PPX_SVD_BLOCK 256
#endif
#ifndef PPX_SVD_DESYNC_NS
#define PPX_SVD_DESYNC_NS 0
#endif
    static constexpr int BLOCK = (M * N >= 36 ? PPX_SVD_BLOCK : 128), MINB = (M * N >= 36 ? PPX_SVD_MINB : 4),
                         MAXW = (M * N > N * N ? M * N : N * N), DESYNC_NS = (M * N >= 36 ? PPX_SVD_DESYNC_NS : 0);
    const double *A;
    double *U;
    double *w;
    double *V;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], v[N * N], w2[N], sv[N];
        io.load(A, a);
        const int shift = pow2_shift(a);
        pow2_rescale(a, shift);
        jacobi_svd_core<M, N, WANT_V>(a, v, w2);
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            // w = w2 / sqrt(w2) and 1/w from one branch-free reciprocal square root (sqrt() and operator/ each carry
            // a slow-path call); exactly zero columns (rank deficiency) stay exactly zero
            const bool pos = w2[j] >= DBL_MIN; // (a denormal squared norm, sigma < 1.5e-154 of the largest entry, counts as 0)
            const double rs = rsqrt_nr(pos ? w2[j] : 1.0);
            const double ws = pos ? w2[j] * rs : (w2[j] < DBL_MIN ? 0.0 : w2[j]); // NaN in, NaN out
            sv[j] = ws;
            if (WANT_U)
            {
                const double inv = pos ? rs : 0.0;
#pragma unroll
                for (int i = 0; i < M; i++)
                    a[i + j * M] *= inv;
            }
        }
        pow2_rescale(sv, -shift);
        if (WANT_U)
            io.store(U, a);
        io.store(w, sv);
        if (WANT_V)
            io.store(V, v);
    }
};

// square systems: LU, or the row-orthogonalising Jacobi solve (no V) when close to singular (square_pinv_solve);
// rectangular ones: column Jacobi with V
template <int N>
__device__ __forceinline__ void solve_scaled(double (&a)[N * N], double (&)[N * N], double (&)[N], const double (&rhs)[N],
                                             double (&sol)[N], bool allow_lu)
{
    square_pinv_solve<N>(a, rhs, sol, allow_lu);
}
template <int M, int N>
__device__ __forceinline__ typename std::enable_if<M != N>::type solve_scaled(double (&a)[M * N], double (&v)[N * N],
                                                                             double (&w2)[N], const double (&rhs)[M],
                                                                             double (&sol)[N], bool)
{
    jacobi_svd_core<M, N, true>(a, v, w2);
    jacobi_svd_solve<M, N>(a, v, w2, rhs, sol);
}

template <int M, int N>
struct OpLinsolveSVD
{
    static constexpr int BLOCK = 128, MINB = (M == N ? 2 : 1), MAXW = M * N;
    const double *A;
    const double *b;
    double *x;
    bool allow_lu;
    // square systems mostly take the LU path and are then bound by their 8 (n^2 + 2n) bytes: inputs staged with cp.async
    // like OpLinsolveLU's
    static constexpr bool PREFETCH = (M == N);
    static constexpr int IN_WIDTH = M * N + M + 2;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<M * N>(A, stage);
        io.template async_load<M>(b, stage + IO::stage_doubles(BLOCK, M * N));
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double a[M * N], rhs[M];
        io.template read_staged<M * N>(stage, a);
        io.template read_staged<M>(stage + IO::stage_doubles(BLOCK, M * N), rhs);
        solve_store(io, a, rhs);
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], rhs[M];
        io.load(A, a);
        io.load(b, rhs);
        solve_store(io, a, rhs);
    }
    template <class IO>
    __device__ __forceinline__ void solve_store(const IO &io, double (&a)[M * N], const double (&rhs)[M]) const
    {
        double v[N * N], w2[N], sol[N];
        const int shift = pow2_shift(a);
        pow2_rescale(a, shift);
        solve_scaled(a, v, w2, rhs, sol, allow_lu);
        pow2_rescale(sol, shift); // (A 2^-shift)^+ b = 2^shift A^+ b
        io.store(x, sol);
    }
};

// BASELINE config 4b at n = 4: 128 B in, 160 B out, ~2k fp64 instructions.
template <int N, bool WANT_Z>
struct OpEigSym
{
    static constexpr int BLOCK = 128, MINB = (N <= 4 ? 3 : 1), MAXW = N * N;
    const double *S;
    double *d;
    double *z;
    int sorted;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double s[N * N], ev[N], vec[N * N];
        io.load(S, s);
        jacobi_eig_sym<N, WANT_Z>(s, sorted != 0, ev, vec);
        io.store(d, ev);
        if (WANT_Z)
            io.store(z, vec);
    }
};

// LU<n> / linsolve<Factorization::LU> / MatrixS::I() / det(), include/linalg.hpp:448-533, 933-965,
// 1176-1187.  HBM-bound like QR.
#ifndef PPX_LU_PREFETCH
#define PPX_LU_PREFETCH 1
#endif
#ifndef PPX_LU_MINB
#define PPX_LU_MINB 2
#endif
template <int N>
struct OpLinsolveLU
{
    static constexpr int BLOCK = 128, MINB = (N >= 6 ? PPX_LU_MINB : 4), MAXW = N * N;
    const double *A;
    const double *b;
    double *x;
    int8_t *status;
    // the next tile's systems are staged (cp.async) behind the elimination of the current one; the stage holds the
    // matrices followed by the right-hand sides
    static constexpr bool PREFETCH = PPX_LU_PREFETCH != 0;
    static constexpr int IN_WIDTH = N * N + N + 2; // >= both parts in either layout (AoS pads each record to odd)
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<N * N>(A, stage);
        io.template async_load<N>(b, stage + IO::stage_doubles(BLOCK, N * N));
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double a[N * N], rhs[N];
        io.template read_staged<N * N>(stage, a);
        io.template read_staged<N>(stage + IO::stage_doubles(BLOCK, N * N), rhs);
        solve_store(io, a, rhs);
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[N * N], rhs[N];
        io.load(A, a);
        io.load(b, rhs);
        solve_store(io, a, rhs);
    }
    template <class IO>
    __device__ __forceinline__ void solve_store(const IO &io, double (&a)[N * N], double (&rhs)[N]) const
    {
        double b0[N];
#pragma unroll
        for (int i = 0; i < N; i++)
            b0[i] = rhs[i];
        bool even;
        const bool singular = lu_factor_solve<N, 1>(a, rhs, even);
#pragma unroll
        for (int i = 0; i < N; i++)
            rhs[i] = singular ? b0[i] : rhs[i]; // linsolve<LU> returns b untouched on SINGULAR (:1182-1186)
        io.store(x, rhs);
        if (status != nullptr && io.valid)
            status[io.i] = singular ? PPX_STATUS_SINGULAR : PPX_STATUS_NORMAL;
    }
};

template <int N>
struct OpInverse
{
#ifndef PPX_INVERSE_PREFETCH
#define PPX_INVERSE_PREFETCH 1
#endif
#ifndef PPX_INVERSE_MINB
#define PPX_INVERSE_MINB 2
#endif
    static constexpr int BLOCK = 128, MINB = (N >= 6 ? PPX_INVERSE_MINB : 4), MAXW = N * N;
    const double *A;
    double *Ai;
    // 8 N^2 bytes each way and little arithmetic in between: the next tile's matrices are staged (cp.async) behind the
    // elimination of the current one
    static constexpr bool PREFETCH = PPX_INVERSE_PREFETCH != 0;
    static constexpr int IN_WIDTH = N * N;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<N * N>(A, stage);
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[N * N];
        io.load(A, a);
        const bool singular = gauss_jordan_inverse<N>(a);
#pragma unroll
        for (int i = 0; i < N * N; i++)
            a[i] = singular ? 0.0 : a[i];
        io.store(Ai, a);
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double a[N * N];
        io.template read_staged<N * N>(stage, a);
        const bool singular = gauss_jordan_inverse<N>(a);
#pragma unroll
        for (int i = 0; i < N * N; i++)
            a[i] = singular ? 0.0 : a[i]; // MatrixS::I() returns {} on SINGULAR (:938-941)
        io.store(Ai, a);
    }
};

template <int N>
struct OpDet
{
    static constexpr int BLOCK = 128, MINB = 4, MAXW = N * N;
    const double *A;
    double *det;
    // Read-dominated (8 N^2 bytes in, 8 out).  On component planes the plain kernel already runs at the HBM peak and
    // staging costs 20 %; on records, where the loads go through the warp's shared-memory tile anyway, staging the
    // next tile with cp.async gains 12 %.
    static constexpr bool PREFETCH = true, PREFETCH_AOS_ONLY = true;
    static constexpr int IN_WIDTH = N * N;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<N * N>(A, stage);
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double a[N * N];
        io.template read_staged<N * N>(stage, a);
        factor_store(io, a);
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[N * N];
        io.load(A, a);
        factor_store(io, a);
    }
    template <class IO>
    __device__ __forceinline__ void factor_store(const IO &io, double (&a)[N * N]) const
    {
        double dummy[N];
        bool even;
        const bool singular = lu_factor_solve<N, 0>(a, dummy, even);
        double D = even ? 1.0 : -1.0; // :959-964
#pragma unroll
        for (int i = 0; i < N; i++)
            D *= a[i + i * N];
        const double out[1] = {singular ? 0.0 : D};
        io.store(det, out);
    }
};

// pinv and the matrix 2-norm are epilogues of the Jacobi SVD (include/linalg.hpp:1211-1227, 926-931)
template <int M, int N>
struct OpPinv
{
    static constexpr int BLOCK = 128, MINB = (M * N >= 36 ? 1 : 3), MAXW = M * N;
    const double *A;
    double *P; // N x M
    bool allow_lu;
    // A square matrix with |A|_inf |A^-1|_inf < 1e10 (measured on the computed inverse) is inverted by Gauss-Jordan
    // elimination: the truncation below cannot act before cond ~ 2e15, so pinv(A) = A^-1 there (the argument of
    // square_pinv_solve; pivot ratios alone would not do, they do not bound the condition number)
    template <class IO, int MM = M>
    __device__ __forceinline__ typename std::enable_if<MM == N, bool>::type fast_square(const IO &io, const double (&a)[M * N],
                                                                                      int shift, double (&g)[N * M]) const
    {
        if (!allow_lu)
            return false;
#pragma unroll
        for (int i = 0; i < N * N; i++)
            g[i] = a[i];
        double pmin, pmax;
        gauss_jordan_inverse<N>(g, pmin, pmax);
        double na = 0.0, ng = 0.0; // inf-norms (max absolute row sum) of A and of the computed inverse
#pragma unroll
        for (int r = 0; r < N; r++)
        {
            double ra = 0.0, rg = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++)
            {
                ra += fabs(a[r + c * N]);
                rg += fabs(g[r + c * N]);
            }
            na = ra > na ? ra : na;
            ng = (rg > ng || rg != rg) ? rg : ng; // a NaN row sum sticks
        }
        const bool easy = na * ng < SQUARE_LU_COND_MAX && pmin > 0.0 && pmax < DBL_MAX; // false for zero, inf, NaN pivots
        pow2_rescale(g, shift);
        if (__all_sync(0xffffffffu, easy))
            io.store(P, g);
        return easy;
    }
    template <class IO, int MM = M>
    __device__ __forceinline__ typename std::enable_if<MM != N, bool>::type fast_square(const IO &, const double (&)[M * N], int,
                                                                                      double (&)[N * M]) const
    {
        return false;
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], v[N * N], w2[N], p[N * M], g[N * M];
        io.load(A, a);
        const int shift = pow2_shift(a);
        pow2_rescale(a, shift);
        const bool easy = fast_square(io, a, shift, g);
        if (M == N && __all_sync(0xffffffffu, easy))
            return; // stored by fast_square
        jacobi_svd_core<M, N, true>(a, v, w2);
        double wmax2 = w2[0];
#pragma unroll
        for (int j = 1; j < N; j++)
            wmax2 = fmax(wmax2, w2[j]);
        const double f = 0.5 * sqrt((double)(M + N + 1)) * EPS_DP;
        const double tsh2 = f * f * wmax2;
#pragma unroll
        for (int i = 0; i < N * M; i++)
            p[i] = 0.0;
        // pinv = V diag(1/w) U^T = sum_j v_j a_j^T / w_j^2 over w_j > tsh (a_j = w_j u_j)
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            const double iw2 = w2[j] > tsh2 ? 1.0 / w2[j] : 0.0;
#pragma unroll
            for (int c = 0; c < M; c++)
            {
                const double t = a[c + j * M] * iw2;
#pragma unroll
                for (int r = 0; r < N; r++)
                    p[r + c * N] = fma(v[r + j * N], t, p[r + c * N]);
            }
        }
        pow2_rescale(p, shift); // (A 2^-shift)^+ = 2^shift A^+
        if (M == N)
        {
#pragma unroll
            for (int i = 0; i < N * M; i++)
                p[i] = easy ? g[i] : p[i];
        }
        io.store(P, p);
    }
};

// sigma_max; a wide matrix (M < N) is decomposed through its transpose (same singular values)
template <int M, int N>
struct OpNorm2
{
    static constexpr int R = (M >= N ? M : N), C = (M >= N ? N : M);
    static constexpr int BLOCK = 128, MINB = (M * N >= 30 ? 2 : 4), MAXW = M * N;
    const double *A;
    double *s;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double a[M * N], t[R * C], v[C * C], w2[C];
        io.load(A, a);
        const int shift = pow2_shift(a);
        pow2_rescale(a, shift);
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
            for (int r = 0; r < R; r++)
                t[r + c * R] = (M >= N) ? a[r + c * M] : a[c + r * M];
        jacobi_svd_core<R, C, false>(t, v, w2);
        double wmax2 = w2[0];
#pragma unroll
        for (int j = 1; j < C; j++)
            wmax2 = fmax(wmax2, w2[j]);
        double out[1] = {sqrt(wmax2)};
        pow2_rescale(out, -shift);
        io.store(s, out);
    }
};

template <int M, int N>
int svd_mn(const double *A, double *U, double *w, double *V, size_t count, int layout, size_t ld, void *stream)
{
    if (U && V)
        return dispatch(layout, OpSVD<M, N, true, true>{A, U, w, V}, count, ld, stream);
    if (U)
        return dispatch(layout, OpSVD<M, N, true, false>{A, U, w, nullptr}, count, ld, stream);
    if (V)
        return dispatch(layout, OpSVD<M, N, false, true>{A, nullptr, w, V}, count, ld, stream);
    return dispatch(layout, OpSVD<M, N, false, false>{A, nullptr, w, nullptr}, count, ld, stream);
}

template <int N>
int eig_n(const double *S, double *d, double *z, int sorted, size_t count, int layout, size_t ld, void *stream)
{
    if (z)
        return dispatch(layout, OpEigSym<N, true>{S, d, z, sorted}, count, ld, stream);
    return dispatch(layout, OpEigSym<N, false>{S, d, nullptr, sorted}, count, ld, stream);
}

} // namespace

#define PPX_SHAPE2(M_, N_, call) \
    if (m == M_ && n == N_)      \
    return call

extern "C"
{

int ppx_cuda_batch_linsolve_qr(int m, int n, const double *A, const double *b, double *x, int8_t *status,
                               size_t count, int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, b, x);
    PPX_SHAPE2(4, 3, dispatch(layout, OpLinsolveQR<4, 3>{A, b, x, status}, count, ld, stream));
    PPX_SHAPE2(3, 3, dispatch(layout, OpLinsolveQR<3, 3>{A, b, x, status}, count, ld, stream));
    PPX_SHAPE2(4, 4, dispatch(layout, OpLinsolveQR<4, 4>{A, b, x, status}, count, ld, stream));
    PPX_SHAPE2(6, 4, dispatch(layout, OpLinsolveQR<6, 4>{A, b, x, status}, count, ld, stream));
    PPX_SHAPE2(6, 6, dispatch(layout, OpLinsolveQR<6, 6>{A, b, x, status}, count, ld, stream));
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count, int layout,
                       size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, w);
    PPX_SHAPE2(6, 6, (svd_mn<6, 6>(A, U, w, V, count, layout, ld, stream)));
    PPX_SHAPE2(4, 3, (svd_mn<4, 3>(A, U, w, V, count, layout, ld, stream)));
    PPX_SHAPE2(3, 3, (svd_mn<3, 3>(A, U, w, V, count, layout, ld, stream)));
    PPX_SHAPE2(4, 4, (svd_mn<4, 4>(A, U, w, V, count, layout, ld, stream)));
    PPX_SHAPE2(5, 4, (svd_mn<5, 4>(A, U, w, V, count, layout, ld, stream)));
    PPX_SHAPE2(6, 4, (svd_mn<6, 4>(A, U, w, V, count, layout, ld, stream)));
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count,
                                int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, b, x);
    PPX_SHAPE2(6, 6, dispatch(layout, OpLinsolveSVD<6, 6>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(4, 3, dispatch(layout, OpLinsolveSVD<4, 3>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(3, 3, dispatch(layout, OpLinsolveSVD<3, 3>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(4, 4, dispatch(layout, OpLinsolveSVD<4, 4>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(5, 4, dispatch(layout, OpLinsolveSVD<5, 4>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(6, 4, dispatch(layout, OpLinsolveSVD<6, 4>{A, b, x, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count, int layout,
                           size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(S, d);
    switch (n)
    {
    case 2: return eig_n<2>(S, d, z, sorted, count, layout, ld, stream);
    case 3: return eig_n<3>(S, d, z, sorted, count, layout, ld, stream);
    case 4: return eig_n<4>(S, d, z, sorted, count, layout, ld, stream);
    case 5: return eig_n<5>(S, d, z, sorted, count, layout, ld, stream);
    case 6: return eig_n<6>(S, d, z, sorted, count, layout, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_linsolve_lu(int n, const double *A, const double *b, double *x, int8_t *status, size_t count,
                               int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, b, x);
    switch (n)
    {
    case 2: return dispatch(layout, OpLinsolveLU<2>{A, b, x, status}, count, ld, stream);
    case 3: return dispatch(layout, OpLinsolveLU<3>{A, b, x, status}, count, ld, stream);
    case 4: return dispatch(layout, OpLinsolveLU<4>{A, b, x, status}, count, ld, stream);
    case 5: return dispatch(layout, OpLinsolveLU<5>{A, b, x, status}, count, ld, stream);
    case 6: return dispatch(layout, OpLinsolveLU<6>{A, b, x, status}, count, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_inverse(int n, const double *A, double *Ai, size_t count, int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, Ai);
    switch (n)
    {
    case 2: return dispatch(layout, OpInverse<2>{A, Ai}, count, ld, stream);
    case 3: return dispatch(layout, OpInverse<3>{A, Ai}, count, ld, stream);
    case 4: return dispatch(layout, OpInverse<4>{A, Ai}, count, ld, stream);
    case 5: return dispatch(layout, OpInverse<5>{A, Ai}, count, ld, stream);
    case 6: return dispatch(layout, OpInverse<6>{A, Ai}, count, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_det(int n, const double *A, double *det, size_t count, int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, det);
    switch (n)
    {
    case 2: return dispatch(layout, OpDet<2>{A, det}, count, ld, stream);
    case 3: return dispatch(layout, OpDet<3>{A, det}, count, ld, stream);
    case 4: return dispatch(layout, OpDet<4>{A, det}, count, ld, stream);
    case 5: return dispatch(layout, OpDet<5>{A, det}, count, ld, stream);
    case 6: return dispatch(layout, OpDet<6>{A, det}, count, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_pinv(int m, int n, const double *A, double *P, size_t count, int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, P);
    PPX_SHAPE2(6, 6, dispatch(layout, OpPinv<6, 6>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(4, 3, dispatch(layout, OpPinv<4, 3>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(3, 3, dispatch(layout, OpPinv<3, 3>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(4, 4, dispatch(layout, OpPinv<4, 4>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(5, 4, dispatch(layout, OpPinv<5, 4>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    PPX_SHAPE2(6, 4, dispatch(layout, OpPinv<6, 4>{A, P, g_ppx_square_lu.load(std::memory_order_relaxed) != 0}, count, ld, stream));
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_norm2(int m, int n, const double *A, double *s, size_t count, int layout, size_t ld, void *stream)
{
    if (count == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, s);
    PPX_SHAPE2(6, 6, dispatch(layout, OpNorm2<6, 6>{A, s}, count, ld, stream));
    PPX_SHAPE2(4, 3, dispatch(layout, OpNorm2<4, 3>{A, s}, count, ld, stream));
    PPX_SHAPE2(3, 3, dispatch(layout, OpNorm2<3, 3>{A, s}, count, ld, stream));
    PPX_SHAPE2(4, 4, dispatch(layout, OpNorm2<4, 4>{A, s}, count, ld, stream));
    PPX_SHAPE2(5, 4, dispatch(layout, OpNorm2<5, 4>{A, s}, count, ld, stream));
    PPX_SHAPE2(4, 5, dispatch(layout, OpNorm2<4, 5>{A, s}, count, ld, stream));
    PPX_SHAPE2(6, 4, dispatch(layout, OpNorm2<6, 4>{A, s}, count, ld, stream));
    return PPX_ERR_BAD_SHAPE;
}

} // extern "C"
