// lie.cu -- batched so3/se3 exp & log, SE3 product / inverse / adjoints (include/liegroup.hpp of the
// reference).  One instance per thread; see ppx_math.cuh for the arithmetic and ppx_io.cuh for the
// HBM <-> register paths.  All of these are a few dozen to ~150 fp64 instructions on 24..288 bytes
// of compulsory traffic per instance: HBM-bound, so the only design rules that matter are
// coalescing, evict-first streaming and enough resident warps to cover DRAM latency.
#include "ppx_launch.cuh"

using namespace ppx_dev;

namespace
{

struct OpSo3Exp
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 9;
    const double *w;
    double *R;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double x[3], r[9];
        io.load(w, x);
        so3_exp(x, r);
        io.store(R, r);
    }
};

struct OpSO3Log
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 9;
    const double *R;
    double *w;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double r[9], x[3];
        io.load(R, r);
        SO3_log(r, x);
        io.store(w, x);
    }
};

struct OpSe3Exp
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 16;
    const double *xi;
    double *T;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double x[6], m[16];
        io.load(xi, x);
        Rigid t;
        se3_exp(x, t);
        rigid_to_record(t, m);
        io.store(T, m);
    }
};

struct OpSE3Log
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 16;
    const double *T;
    double *xi;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double m[16], x[6];
        io.load(T, m);
        Rigid t;
        rigid_from_record(m, t);
        SE3_log(t, x);
        io.store(xi, x);
    }
};

// BASELINE config 1: xi -> exp -> log -> xi', the SE3 never leaves registers (96 B/instance instead of 352).
struct OpSe3ExpLog
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 6;
    const double *xi;
    double *xo;
    static constexpr bool PREFETCH = true; // 82.6 % -> 88 % of the copy roofline (93 % with the branch-free 1/phi)
    static constexpr int IN_WIDTH = 6;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<6>(xi, stage);
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double x[6], y[6];
        io.template read_staged<6>(stage, x);
        Rigid t;
        se3_exp(x, t);
        SE3_log(t, y);
        io.store(xo, y);
    }
};

struct OpSE3Mul
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 16;
    const double *A, *B;
    double *C;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double m[16];
        Rigid a, b, c;
        io.load(A, m);
        rigid_from_record(m, a);
        io.load(B, m);
        rigid_from_record(m, b);
        rigid_mul(a, b, c);
        rigid_to_record(c, m);
        io.store(C, m);
    }
};

struct OpSE3Inv
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 16;
    const double *T;
    double *Ti;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double m[16];
        Rigid a, b;
        io.load(T, m);
        rigid_from_record(m, a);
        rigid_inv(a, b);
        rigid_to_record(b, m);
        io.store(Ti, m);
    }
};

struct OpSE3Adt
{
    static constexpr int BLOCK = 128, MINB = 4, MAXW = 36;
    const double *T;
    double *Ad;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double m[16], ad[36];
        Rigid a;
        io.load(T, m);
        rigid_from_record(m, a);
        rigid_adjoint_record(a, ad);
        io.store(Ad, ad);
    }
};

struct OpSe3Adt
{
    static constexpr int BLOCK = 128, MINB = 4, MAXW = 36;
    const double *xi;
    double *ad;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double x[6], m[36];
        io.load(xi, x);
        se3_adt_record(x, m);
        io.store(ad, m);
    }
};

struct OpSo3Jac
{
    static constexpr int BLOCK = 256, MINB = 3, MAXW = 9;
    const double *w;
    double *J;
    int which;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double x[3], j[9];
        io.load(w, x);
        so3_jac(which, x, j);
        io.store(J, j);
    }
};

} // namespace

extern "C"
{

int ppx_cuda_batch_so3_exp(const double *w, double *R, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(w, R);
    return dispatch(layout, OpSo3Exp{w, R}, n, ld, stream);
}

int ppx_cuda_batch_SO3_log(const double *R, double *w, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(w, R);
    return dispatch(layout, OpSO3Log{R, w}, n, ld, stream);
}

int ppx_cuda_batch_se3_exp(const double *xi, double *T, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(xi, T);
    return dispatch(layout, OpSe3Exp{xi, T}, n, ld, stream);
}

int ppx_cuda_batch_SE3_log(const double *T, double *xi, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(xi, T);
    return dispatch(layout, OpSE3Log{T, xi}, n, ld, stream);
}

int ppx_cuda_batch_se3_exp_log(const double *xi, double *xi_out, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(xi, xi_out);
    return dispatch(layout, OpSe3ExpLog{xi, xi_out}, n, ld, stream);
}

int ppx_cuda_batch_SE3_mul(const double *A, const double *B, double *C, size_t n, int layout, size_t ld,
                           void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(A, B, C);
    return dispatch(layout, OpSE3Mul{A, B, C}, n, ld, stream);
}

int ppx_cuda_batch_SE3_inv(const double *T, double *Ti, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(T, Ti);
    return dispatch(layout, OpSE3Inv{T, Ti}, n, ld, stream);
}

int ppx_cuda_batch_SE3_Adt(const double *T, double *Ad, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(T, Ad);
    return dispatch(layout, OpSE3Adt{T, Ad}, n, ld, stream);
}

int ppx_cuda_batch_se3_adt(const double *xi, double *ad, size_t n, int layout, size_t ld, void *stream)
{
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(xi, ad);
    return dispatch(layout, OpSe3Adt{xi, ad}, n, ld, stream);
}

int ppx_cuda_batch_so3_jac(int which, const double *w, double *J, size_t n, int layout, size_t ld, void *stream)
{
    if (which < 0 || which > 3)
        return PPX_ERR_BAD_ARGUMENT;
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(w, J);
    return dispatch(layout, OpSo3Jac{w, J, which}, n, ld, stream);
}

} // extern "C"
