// math_host.cpp -- TEST HARNESS: the per-lane solver math of matrix_b200/csrc/ppx_math.cuh compiled for the host.
//
// The Jacobi SVD (with its look-ahead), the Jacobi eigensolver, the LU / Gauss-Jordan routines and the square
// pseudo-inverse solve with its LU fast path and SVD fallback are plain C++ apart from the CUDA qualifiers; here the
// very code the GPU runs per lane is called on arrays of dense reference records so that tests/test_math_host.py can
// hold it against numpy and the oracle WITHOUT a GPU.  Not a product path (the product has no CPU path): nothing under
// matrix_b200/ or include/ builds or loads this file.
//
//   g++ -O2 -std=c++17 -shared -fPIC -ffp-contract=off -Wno-unknown-pragmas tests/cpp/math_host.cpp -o math_host.so
#include "cuda_host_shims.h"

#include "../../include/ppx_cuda.h"
#include "../../matrix_b200/csrc/ppx_math.cuh"

using namespace ppx_dev;

template <int M, int N>
static void svd_mn(const double *A, double *US, double *w2, double *V, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double a[M * N], v[N * N], w[N];
        std::memcpy(a, A + M * N * i, sizeof(a));
        jacobi_svd_core<M, N, true>(a, v, w);
        std::memcpy(US + M * N * i, a, sizeof(a));
        std::memcpy(V + N * N * i, v, sizeof(v));
        std::memcpy(w2 + N * i, w, sizeof(w));
    }
}

extern "C"
{

// A (m x n records) -> US = U diag(w) (m x n), w2 = w^2 (n), V (n x n); shapes 6x6 and 4x3
int mh_svd(int m, int n_, const double *A, double *US, double *w2, double *V, size_t n)
{
    if (m == 6 && n_ == 6)
        svd_mn<6, 6>(A, US, w2, V, n);
    else if (m == 4 && n_ == 3)
        svd_mn<4, 3>(A, US, w2, V, n);
    else
        return -1;
    return 0;
}

// symmetric 4x4 / 6x6 -> eigenvalues (ascending) and eigenvectors in columns
int mh_eig(int n_, const double *S, double *d, double *Z, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        if (n_ == 4)
        {
            double s[16], ev[4], z[16];
            std::memcpy(s, S + 16 * i, sizeof(s));
            jacobi_eig_sym<4, true>(s, true, ev, z);
            std::memcpy(d + 4 * i, ev, sizeof(ev));
            std::memcpy(Z + 16 * i, z, sizeof(z));
        }
        else if (n_ == 6)
        {
            double s[36], ev[6], z[36];
            std::memcpy(s, S + 36 * i, sizeof(s));
            jacobi_eig_sym<6, true>(s, true, ev, z);
            std::memcpy(d + 6 * i, ev, sizeof(ev));
            std::memcpy(Z + 36 * i, z, sizeof(z));
        }
        else
            return -1;
    }
    return 0;
}

// 6x6: inverse by in-place Gauss-Jordan (singular flag per instance), LU solve, determinant sign bookkeeping
int mh_inverse6(const double *A, double *Ai, signed char *singular, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double a[36];
        std::memcpy(a, A + 36 * i, sizeof(a));
        singular[i] = gauss_jordan_inverse<6>(a) ? 1 : 0;
        std::memcpy(Ai + 36 * i, a, sizeof(a));
    }
    return 0;
}
int mh_lu_solve6(const double *A, const double *b, double *x, signed char *singular, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double a[36], r[6];
        std::memcpy(a, A + 36 * i, sizeof(a));
        std::memcpy(r, b + 6 * i, sizeof(r));
        bool even;
        singular[i] = lu_factor_solve<6, 1>(a, r, even) ? 1 : 0;
        std::memcpy(x + 6 * i, r, sizeof(r));
    }
    return 0;
}

// linsolve<SVD> on a square 6x6 system: LU fast path (allow_lu != 0) with the Jacobi SVD fallback, or the fallback only
int mh_square_pinv_solve6(const double *A, const double *b, double *x, int allow_lu, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double a[36], r[6], sol[6];
        std::memcpy(a, A + 36 * i, sizeof(a));
        std::memcpy(r, b + 6 * i, sizeof(r));
        square_pinv_solve<6>(a, r, sol, allow_lu != 0);
        std::memcpy(x + 6 * i, sol, sizeof(sol));
    }
    return 0;
}

// ---- Lie maps and kinematics (ppx_math.cuh: se3_exp, SE3_log, so3_exp, SO3_log, poe_fk_jacobian) ----
int mh_se3_exp(const double *xi, double *T, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double x[6], m[16];
        std::memcpy(x, xi + 6 * i, sizeof(x));
        Rigid t;
        se3_exp(x, t);
        rigid_to_record(t, m);
        std::memcpy(T + 16 * i, m, sizeof(m));
    }
    return 0;
}
int mh_SE3_log(const double *T, double *xi, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double x[6], m[16];
        std::memcpy(m, T + 16 * i, sizeof(m));
        Rigid t;
        rigid_from_record(m, t);
        SE3_log(t, x);
        std::memcpy(xi + 6 * i, x, sizeof(x));
    }
    return 0;
}
int mh_so3_exp(const double *w, double *R, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double x[3], r[9];
        std::memcpy(x, w + 3 * i, sizeof(x));
        so3_exp(x, r);
        std::memcpy(R + 9 * i, r, sizeof(r));
    }
    return 0;
}
int mh_SO3_log(const double *R, double *w, size_t n)
{
    for (size_t i = 0; i < n; i++)
    {
        double x[3], r[9];
        std::memcpy(r, R + 9 * i, sizeof(r));
        SO3_log(r, x);
        std::memcpy(w + 3 * i, x, sizeof(x));
    }
    return 0;
}
// 6-joint chain: T (16) and Js (36) per instance
int mh_fk_jacobian6(const double *screws, const double *home, const double *q, double *T, double *Js, size_t n)
{
    ChainConst<6> ch;
    fold_chain<6>(screws, home, ch);
    for (size_t i = 0; i < n; i++)
    {
        double qi[6], m[16], js[36];
        std::memcpy(qi, q + 6 * i, sizeof(qi));
        Rigid t;
        poe_fk_jacobian<6, true, true>(ch, qi, t, [&](int col, const double(&v)[6]) {
            for (int k = 0; k < 6; k++)
                js[6 * col + k] = v[k];
        });
        rigid_to_record(t, m);
        std::memcpy(T + 16 * i, m, sizeof(m));
        std::memcpy(Js + 36 * i, js, sizeof(js));
    }
    return 0;
}

} // extern "C"
