// ref_shim.cpp -- extern "C" batch loops over the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY (see ppx_oracle.h).  This file contains no numerical
// code of its own: every function below is an OpenMP loop whose body is the
// reference call chain named in the comment, on values reinterpreted from the
// caller's dense arrays.  It is compiled by oracle/Makefile with
// -I/root/reference/include -I/root/reference/demo, i.e. from the reference
// sources where they lie; nothing from /root/reference is copied into the repo.
// The output, oracle/_ref/libppx_ref.so, is git-ignored and travels to the GPU
// box as a prebuilt file.
//
// Pinned build: -std=c++14 -O2 -DNDEBUG, no PPX_USE_AVX,
// no -ffast-math.  Two variants are produced: libppx_ref.so (x86-64-v3, GCC's
// default FMA contraction: the reference's Release build, used as the CPU
// baseline) and libppx_ref_strict.so (-ffp-contract=off, used to pin the C
// restatement bit for bit).

#include <string>
#include <algorithm>
#include <cstring>
#include <omp.h>

#include "liegroup.hpp"
#include "robotics.hpp" // brings `using namespace ppx;`

#include "ppx_oracle.h"

namespace
{
    inline int threads_or_all(int nthreads)
    {
        return nthreads > 0 ? nthreads : omp_get_max_threads();
    }

    template <size_t M, size_t N>
    inline MatrixS<M, N> load(const double *p)
    {
        MatrixS<M, N> r;
        std::memcpy(r.data(), p, sizeof(double) * M * N);
        return r;
    }

    template <size_t M, size_t N>
    inline void store(double *p, const MatrixS<M, N> &v)
    {
        std::memcpy(p, v.data(), sizeof(double) * M * N);
    }

    template <size_t N>
    kinematics<N> make_chain(const double *screws, const double *home)
    {
        kinematics<N> kin;
        for (size_t i = 0; i < N; i++)
        {
            kin.JointList()[i] = joint("J" + std::to_string(i), load<6, 1>(screws + 6 * i), SE3());
        }
        kin.JointList()[N - 1].pose = SE3(load<4, 4>(home));
        return kin;
    }

    template <size_t N>
    int fk_jacobian(const double *screws, const double *home, const double *q, double *T, double *Js,
                    size_t count, int nthreads)
    {
        const kinematics<N> chain = make_chain<N>(screws, home);
#pragma omp parallel num_threads(threads_or_all(nthreads))
        {
        kinematics<N> kin = chain; // one private copy per thread (jacobiSpace is non-const in the reference)
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)count; i++)
        {
            const auto qi = load<N, 1>(q + N * i);
            if (T)
            {
                store(T + 16 * i, MatrixS<4, 4>(kin.forwardSpace(qi))); // demo/robotics.hpp:83-92
            }
            if (Js)
            {
                store(Js + 6 * N * i, kin.jacobiSpace(qi)); // demo/robotics.hpp:94-106
            }
        }
        }
        return 0;
    }

    template <size_t N>
    int ik_step(const double *screws, const double *home, const double *q, const double *Tt, double *q_out,
                double *Vs_out, size_t count, int nthreads)
    {
        const kinematics<N> chain = make_chain<N>(screws, home);
#pragma omp parallel num_threads(threads_or_all(nthreads))
        {
        kinematics<N> kin = chain; // one private copy per thread
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)count; i++)
        {
            auto qi = load<N, 1>(q + N * i);
            const SE3 pose(load<4, 4>(Tt + 16 * i));
            // test/lie_test.cpp:104-110
            SE3 Tsb = kin.forwardSpace(qi);
            se3 Vs = Tsb.Adt() * (Tsb.I() * pose).log();
            qi += linsolve<Factorization::SVD>(kin.jacobiSpace(qi), Vs).x;
            store(q_out + N * i, qi);
            if (Vs_out)
            {
                store(Vs_out + 6 * i, Vs);
            }
        }
        }
        return 0;
    }

    template <size_t N>
    int ik_loop(const double *screws, const double *home, const double *q0, const double *Tt, int max_iter,
                double *q_out, int *iters, size_t count, int nthreads)
    {
        const kinematics<N> chain = make_chain<N>(screws, home);
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            kinematics<N> kin = chain;
            auto init = load<N, 1>(q0 + N * i);
            const SE3 pose(load<4, 4>(Tt + 16 * i));
            // test/lie_test.cpp:96-113 (printf removed, 20 -> max_iter)
            bool converged = false;
            int iter = 0;
            while (!converged && iter < max_iter)
            {
                SE3 Tsb = kin.forwardSpace(init);
                se3 Vs = Tsb.Adt() * (Tsb.I() * pose).log();
                auto err_w = norm2(Vs._1());
                auto err_v = norm2(Vs._2());
                converged = err_w < EPS_SP && err_v < EPS_SP;
                init += linsolve<Factorization::SVD>(kin.jacobiSpace(init), Vs).x;
                ++iter;
            }
            store(q_out + N * i, init);
            if (iters)
            {
                iters[i] = iter;
            }
        }
        return 0;
    }

    template <size_t N>
    int ik_codo(const double *screws, const double *home, const double *lo, const double *hi, const double *Tt,
                const double *init, double *q_out, double *x_raw, double *fnorm, signed char *status, size_t count,
                int nthreads)
    {
        kinematics<N> chain = make_chain<N>(screws, home);
        for (size_t i = 0; i < N; i++)
        {
            chain.JointList()[i].range = {lo[i], hi[i]};
        }
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            kinematics<N> kin = chain;
            const auto q0 = load<N, 1>(init + N * i);
            const SE3 pose(load<4, 4>(Tt + 16 * i));
            store(q_out + N * i, kin.inverseSpace(pose, q0)); // demo/robotics.hpp:126-152
            if (x_raw || fnorm || status)
            {
                // the same optimiser call as inverseSpace makes (:128-150), to expose OptResult's y and s as well
                auto fx = [&kin, &pose](const MatrixS<N, 1> &q)
                { return (kin.forwardSpace(q) * pose.I()).log(); };
                auto dfx = [&kin, &pose](const MatrixS<N, 1> &q, const se3 &)
                {
                    auto Tsb = kin.forwardSpace(q) * pose.I();
                    auto JpI = eye<6>();
                    JpI.template sub<3, 3, false>(3, 0) = -1 * hat(Tsb.Pos());
                    return JpI * kin.jacobiSpace(q);
                };
                details::CoDo<N, 6> codo(fx, dfx, load<N, 1>(lo), load<N, 1>(hi));
                auto r = codo(q0);
                if (x_raw)
                {
                    store(x_raw + N * i, r.x);
                }
                if (fnorm)
                {
                    fnorm[i] = r.y;
                }
                if (status)
                {
                    status[i] = (signed char)r.s;
                }
            }
        }
        return 0;
    }

    template <size_t M, size_t N>
    int qr_solve(const double *A, const double *b, double *x, signed char *status, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            auto r = linsolve<Factorization::QR>(load<M, N>(A + M * N * i), load<M, 1>(b + M * i)); // linalg.hpp:1189-1200
            store(x + N * i, r.x);
            if (status)
            {
                status[i] = (signed char)r.s;
            }
        }
        return 0;
    }

    template <size_t N>
    int lu_solve(const double *A, const double *b, double *x, signed char *status, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            auto r = linsolve<Factorization::LU>(load<N, N>(A + N * N * i), load<N, 1>(b + N * i)); // linalg.hpp:1176-1187
            store(x + N * i, r.x);
            if (status)
            {
                status[i] = (signed char)r.s;
            }
        }
        return 0;
    }

    template <size_t N>
    int inv_det(const double *A, double *Ai, double *det, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            const auto a = load<N, N>(A + N * N * i);
            if (Ai)
            {
                store(Ai + N * N * i, a.I()); // linalg.hpp:933-948
            }
            if (det)
            {
                det[i] = a.det(); // linalg.hpp:950-965
            }
        }
        return 0;
    }

    template <size_t M, size_t N>
    int svd_full(const double *A, double *U, double *w, double *V, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            SVD<M, N> s(load<M, N>(A + M * N * i)); // linalg.hpp:605-884
            if (U)
            {
                store(U + M * N * i, s.u);
            }
            store(w + N * i, s.w);
            if (V)
            {
                store(V + N * N * i, s.v);
            }
        }
        return 0;
    }

    template <size_t M, size_t N>
    int svd_solve(const double *A, const double *b, double *x, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            auto r = linsolve<Factorization::SVD>(load<M, N>(A + M * N * i), load<M, 1>(b + M * i)); // linalg.hpp:1202-1209
            store(x + N * i, r.x);
        }
        return 0;
    }

    template <size_t M, size_t N>
    int pinv_norm(const double *A, double *P, double *s, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            const auto a = load<M, N>(A + M * N * i);
            if (P)
            {
                store(P + M * N * i, pinv(a)); // linalg.hpp:1211-1227
            }
            if (s)
            {
                s[i] = norm2(a); // linalg.hpp:926-931
            }
        }
        return 0;
    }

    template <size_t N>
    int eig(const double *S, double *d, double *z, int sorted, size_t count, int nthreads)
    {
#pragma omp parallel for schedule(static) num_threads(threads_or_all(nthreads))
        for (long long i = 0; i < (long long)count; i++)
        {
            EigenValue<N> e(load<N, N>(S + N * N * i), sorted != 0); // linalg.hpp:1154-1163
            store(d + N * i, e.d);
            if (z)
            {
                store(z + N * N * i, e.z);
            }
        }
        return 0;
    }
}

#define PPXO_LOOP(...)                                                             \
    _Pragma("omp parallel for schedule(static) num_threads(threads_or_all(nthreads))") \
    for (long long i = 0; i < (long long)count; i++) { __VA_ARGS__ }               \
    return 0;

extern "C"
{
    const char *ppxo_kind(void) { return "reference"; }
    int ppxo_max_threads(void) { return omp_get_max_threads(); }

    int ppxo_so3_exp(const double *w, double *R, size_t count, int nthreads)
    {
        PPXO_LOOP(store(R + 9 * i, MatrixS<3, 3>(load<3, 1>(w + 3 * i).exp()));) // liegroup.hpp:116-133
    }
    int ppxo_SO3_log(const double *R, double *w, size_t count, int nthreads)
    {
        PPXO_LOOP(store(w + 3 * i, SO3(load<3, 3>(R + 9 * i)).log());) // liegroup.hpp:60-92
    }
    int ppxo_se3_exp(const double *xi, double *T, size_t count, int nthreads)
    {
        PPXO_LOOP(store(T + 16 * i, MatrixS<4, 4>(load<6, 1>(xi + 6 * i).exp()));) // liegroup.hpp:253-269
    }
    int ppxo_SE3_log(const double *T, double *xi, size_t count, int nthreads)
    {
        PPXO_LOOP(store(xi + 6 * i, SE3(load<4, 4>(T + 16 * i)).log());) // liegroup.hpp:231-250
    }
    int ppxo_se3_exp_log(const double *xi, double *xo, size_t count, int nthreads)
    {
        PPXO_LOOP(store(xo + 6 * i, load<6, 1>(xi + 6 * i).exp().log());)
    }
#define PPXO_MATMUL(M_, N_, L_)                                                                                  \
    if (m == M_ && n == N_ && l == L_)                                                                           \
    {                                                                                                            \
        PPXO_LOOP(store(C + M_ * L_ * i, load<M_, N_>(A + M_ * N_ * i) * load<N_, L_>(B + N_ * L_ * i));)        \
    }
    int ppxo_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count, int nthreads)
    {
        // MatrixS<M,N>::operator*(const MatrixS<N,L> &), include/matrixs.hpp:596-640
        PPXO_MATMUL(3, 3, 3)
        PPXO_MATMUL(4, 4, 4)
        PPXO_MATMUL(6, 6, 6)
        PPXO_MATMUL(6, 6, 1)
        PPXO_MATMUL(4, 4, 1)
        PPXO_MATMUL(3, 3, 1)
        PPXO_MATMUL(4, 3, 3)
        PPXO_MATMUL(5, 4, 2)
        PPXO_MATMUL(2, 7, 3)
        return -1;
    }
    int ppxo_SE3_mul(const double *A, const double *B, double *C, size_t count, int nthreads)
    {
        PPXO_LOOP(store(C + 16 * i, MatrixS<4, 4>(SE3(load<4, 4>(A + 16 * i)) * SE3(load<4, 4>(B + 16 * i))));) // liegroup.hpp:204-207
    }
    int ppxo_SE3_inv(const double *T, double *Ti, size_t count, int nthreads)
    {
        PPXO_LOOP(store(Ti + 16 * i, MatrixS<4, 4>(SE3(load<4, 4>(T + 16 * i)).I()));) // liegroup.hpp:227-230
    }
    int ppxo_SE3_Adt(const double *T, double *Ad, size_t count, int nthreads)
    {
        PPXO_LOOP(store(Ad + 36 * i, SE3(load<4, 4>(T + 16 * i)).Adt());) // liegroup.hpp:217-226
    }
    int ppxo_se3_adt(const double *xi, double *ad, size_t count, int nthreads)
    {
        // liegroup.hpp:271-280 leaves the (0,3) block to MatrixS's zero-initialising default ctor.
        PPXO_LOOP(store(ad + 36 * i, load<6, 1>(xi + 6 * i).adt());)
    }
    int ppxo_so3_jac(int which, const double *w, double *J, size_t count, int nthreads)
    {
        if (which < 0 || which > 3)
        {
            return -1;
        }
        PPXO_LOOP(const auto v = load<3, 1>(w + 3 * i);
                  store(J + 9 * i, which == 0 ? v.ljac() : which == 1 ? v.ljacinv()
                                                   : which == 2   ? v.rjac()
                                                                  : v.rjacinv());) // liegroup.hpp:142-184
    }

    int ppxo_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q, double *T,
                             double *Js, size_t count, int nthreads)
    {
        switch (njoints)
        {
        case 2: return fk_jacobian<2>(screws, home, q, T, Js, count, nthreads);
        case 3: return fk_jacobian<3>(screws, home, q, T, Js, count, nthreads);
        case 4: return fk_jacobian<4>(screws, home, q, T, Js, count, nthreads);
        case 5: return fk_jacobian<5>(screws, home, q, T, Js, count, nthreads);
        case 6: return fk_jacobian<6>(screws, home, q, T, Js, count, nthreads);
        case 7: return fk_jacobian<7>(screws, home, q, T, Js, count, nthreads);
        default: return -1;
        }
    }

    int ppxo_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                            const double *Ttarget, double *q_out, double *Vs, size_t count, int nthreads)
    {
        switch (njoints)
        {
        case 4: return ik_step<4>(screws, home, q, Ttarget, q_out, Vs, count, nthreads);
        case 5: return ik_step<5>(screws, home, q, Ttarget, q_out, Vs, count, nthreads);
        case 6: return ik_step<6>(screws, home, q, Ttarget, q_out, Vs, count, nthreads);
        default: return -1; // linsolve<SVD> on a 6 x N Jacobian needs N <= 6 (linalg.hpp:1208)
        }
    }

    int ppxo_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                      const double *Ttarget, int max_iter, double *q_out, int *iters, size_t count, int nthreads)
    {
        switch (njoints)
        {
        case 6: return ik_loop<6>(screws, home, q0, Ttarget, max_iter, q_out, iters, count, nthreads);
        default: return -1;
        }
    }

    int ppxo_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                     const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                     signed char *status, size_t count, int nthreads)
    {
        if (njoints != 6)
        {
            return -1; // CoDo<N,6> solves a square system through linsolve<LU> (optimization.hpp:608)
        }
        return ik_codo<6>(screws, home, lo, hi, Ttarget, init, q_out, x_raw, fnorm, status, count, nthreads);
    }

#define SHAPE2(fn, M_, N_, ...) \
    if (m == M_ && n == N_)     \
    return fn<M_, N_>(__VA_ARGS__)
#define SHAPE1(fn, N_, ...) \
    if (n == N_)            \
    return fn<N_>(__VA_ARGS__)

    int ppxo_linsolve_qr(int m, int n, const double *A, const double *b, double *x, signed char *status,
                         size_t count, int nthreads)
    {
        SHAPE2(qr_solve, 4, 3, A, b, x, status, count, nthreads);
        SHAPE2(qr_solve, 3, 3, A, b, x, status, count, nthreads);
        SHAPE2(qr_solve, 4, 4, A, b, x, status, count, nthreads);
        SHAPE2(qr_solve, 6, 4, A, b, x, status, count, nthreads);
        SHAPE2(qr_solve, 6, 6, A, b, x, status, count, nthreads);
        SHAPE2(qr_solve, 8, 5, A, b, x, status, count, nthreads);
        return -1;
    }
    int ppxo_linsolve_lu(int n, const double *A, const double *b, double *x, signed char *status, size_t count,
                         int nthreads)
    {
        SHAPE1(lu_solve, 2, A, b, x, status, count, nthreads);
        SHAPE1(lu_solve, 3, A, b, x, status, count, nthreads);
        SHAPE1(lu_solve, 4, A, b, x, status, count, nthreads);
        SHAPE1(lu_solve, 5, A, b, x, status, count, nthreads);
        SHAPE1(lu_solve, 6, A, b, x, status, count, nthreads);
        return -1;
    }
    int ppxo_inverse(int n, const double *A, double *Ai, size_t count, int nthreads)
    {
        SHAPE1(inv_det, 2, A, Ai, nullptr, count, nthreads);
        SHAPE1(inv_det, 3, A, Ai, nullptr, count, nthreads);
        SHAPE1(inv_det, 4, A, Ai, nullptr, count, nthreads);
        SHAPE1(inv_det, 5, A, Ai, nullptr, count, nthreads);
        SHAPE1(inv_det, 6, A, Ai, nullptr, count, nthreads);
        return -1;
    }
    int ppxo_det(int n, const double *A, double *det, size_t count, int nthreads)
    {
        SHAPE1(inv_det, 2, A, nullptr, det, count, nthreads);
        SHAPE1(inv_det, 3, A, nullptr, det, count, nthreads);
        SHAPE1(inv_det, 4, A, nullptr, det, count, nthreads);
        SHAPE1(inv_det, 5, A, nullptr, det, count, nthreads);
        SHAPE1(inv_det, 6, A, nullptr, det, count, nthreads);
        return -1;
    }
    int ppxo_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count, int nthreads)
    {
        SHAPE2(svd_full, 4, 3, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 6, 6, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 3, 3, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 4, 4, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 5, 4, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 4, 5, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 6, 4, A, U, w, V, count, nthreads);
        SHAPE2(svd_full, 6, 5, A, U, w, V, count, nthreads);
        return -1;
    }
    int ppxo_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count, int nthreads)
    {
        SHAPE2(svd_solve, 4, 3, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 6, 6, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 3, 3, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 4, 4, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 5, 4, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 6, 4, A, b, x, count, nthreads);
        SHAPE2(svd_solve, 6, 5, A, b, x, count, nthreads);
        return -1;
    }
    int ppxo_pinv(int m, int n, const double *A, double *P, size_t count, int nthreads)
    {
        SHAPE2(pinv_norm, 4, 3, A, P, nullptr, count, nthreads);
        SHAPE2(pinv_norm, 6, 6, A, P, nullptr, count, nthreads);
        SHAPE2(pinv_norm, 3, 3, A, P, nullptr, count, nthreads);
        SHAPE2(pinv_norm, 4, 4, A, P, nullptr, count, nthreads);
        SHAPE2(pinv_norm, 5, 4, A, P, nullptr, count, nthreads);
        return -1;
    }
    int ppxo_norm2(int m, int n, const double *A, double *s, size_t count, int nthreads)
    {
        SHAPE2(pinv_norm, 4, 3, A, nullptr, s, count, nthreads);
        SHAPE2(pinv_norm, 6, 6, A, nullptr, s, count, nthreads);
        SHAPE2(pinv_norm, 3, 3, A, nullptr, s, count, nthreads);
        SHAPE2(pinv_norm, 4, 4, A, nullptr, s, count, nthreads);
        SHAPE2(pinv_norm, 5, 4, A, nullptr, s, count, nthreads);
        SHAPE2(pinv_norm, 4, 5, A, nullptr, s, count, nthreads);
        return -1;
    }
    int ppxo_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count, int nthreads)
    {
        SHAPE1(eig, 2, S, d, z, sorted, count, nthreads);
        SHAPE1(eig, 3, S, d, z, sorted, count, nthreads);
        SHAPE1(eig, 4, S, d, z, sorted, count, nthreads);
        SHAPE1(eig, 5, S, d, z, sorted, count, nthreads);
        SHAPE1(eig, 6, S, d, z, sorted, count, nthreads);
        return -1;
    }
}
