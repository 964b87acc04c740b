// kinematics.cu -- product-of-exponentials forward kinematics + space Jacobian and the IK Newton
// step built from them (demo/robotics.hpp:83-106 and test/lie_test.cpp:96-113 of the reference).
//
// The chain (screw axes + home pose) is the same for every instance of a call, so the host folds
// everything that depends on it alone into a ChainConst that travels as a __grid_constant__ kernel
// parameter (constant bank: free FMA operands, no registers).  Per instance the kernel then needs
// only NJ sincos, NJ-1 rigid products shared between the pose and the Jacobian columns, and NJ-1
// structured adjoint applications: no division, no square root, no 4x4 matrix powers.
#include <cmath>

#include "ppx_launch.cuh"

using namespace ppx_dev;

namespace
{

template <int NJ>
void fold_chain(const double *screws, const double *home, ChainConst<NJ> &ch)
{
    for (int i = 0; i < NJ; i++)
    {
        JointConst &j = ch.joint[i];
        const double *s = screws + 6 * i;
        for (int k = 0; k < 3; k++)
        {
            j.w[k] = s[k];
            j.v[k] = s[3 + k];
        }
        j.wv[0] = j.w[1] * j.v[2] - j.w[2] * j.v[1];
        j.wv[1] = j.w[2] * j.v[0] - j.w[0] * j.v[2];
        j.wv[2] = j.w[0] * j.v[1] - j.w[1] * j.v[0];
        j.wwv[0] = j.w[1] * j.wv[2] - j.w[2] * j.wv[1];
        j.wwv[1] = j.w[2] * j.wv[0] - j.w[0] * j.wv[2];
        j.wwv[2] = j.w[0] * j.wv[1] - j.w[1] * j.wv[0];
        j.wn = std::sqrt(j.w[0] * j.w[0] + j.w[1] * j.w[1] + j.w[2] * j.w[2]);
        j.iw = j.wn > 0.0 ? 1.0 / j.wn : 0.0;
        j.iw2 = j.iw * j.iw;
        j.iw3 = j.iw2 * j.iw;
    }
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            ch.home_r[r + 3 * c] = home[r + 4 * c];
    ch.home_p[0] = home[12], ch.home_p[1] = home[13], ch.home_p[2] = home[14];
}

// BASELINE config 2.  48 B in, 128 + 288 B out per instance at NJ = 6: write-dominated and HBM-bound.
template <int NJ, bool WANT_T, bool WANT_J>
struct OpFkJacobian
{
    static constexpr int BLOCK = 128, MINB = 4, MAXW = (WANT_J ? 6 * NJ : 16) > 16 ? (WANT_J ? 6 * NJ : 16) : 16;
    ChainConst<NJ> ch;
    const double *q;
    double *T;
    double *Js;
    static constexpr bool PREFETCH = true;
    static constexpr int IN_WIDTH = NJ;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<NJ>(q, stage);
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double qi[NJ];
        io.template read_staged<NJ>(stage, qi);
        Rigid t;
        // Jacobian columns stream out (SoA: straight to HBM; AoS: into the warp's staging tile) as they are
        // produced, so the 6 x NJ matrix never occupies registers
        poe_fk_jacobian<NJ, WANT_T, WANT_J>(ch, qi, t, [&](int col, const double(&v)[6]) {
            io.template put<6 * NJ, 6>(Js, 6 * col, v);
        });
        if (WANT_J)
            io.template flush<6 * NJ>(Js);
        if (WANT_T)
        {
            double m[16];
            rigid_to_record(t, m);
            io.store(T, m);
        }
    }
};

// Vs = Ad(T(q)) log(T(q)^-1 Tt),  dq = pinv(Js(q)) Vs  (test/lie_test.cpp:104-110)
template <int NJ>
__device__ __forceinline__ void ik_newton_update(const ChainConst<NJ> &ch, const Rigid &target, double (&q)[NJ],
                                                 double (&vs)[6])
{
    Rigid t, ti, x;
    double js[6 * NJ];
    poe_fk_jacobian<NJ, true, true>(ch, q, t, [&](int col, const double(&v)[6]) {
#pragma unroll
        for (int k = 0; k < 6; k++)
            js[6 * col + k] = v[k];
    });
    rigid_inv(t, ti);
    rigid_mul(ti, target, x);
    double lg[6];
    SE3_log(x, lg);
    const double lw[3] = {lg[0], lg[1], lg[2]}, lv[3] = {lg[3], lg[4], lg[5]};
    rigid_adjoint_apply(t, lw, lv, vs);
    double V[NJ * NJ], w2[NJ], dq[NJ];
    jacobi_svd_core<6, NJ, true>(js, V, w2);
    jacobi_svd_solve<6, NJ>(js, V, w2, vs, dq);
#pragma unroll
    for (int i = 0; i < NJ; i++)
        q[i] += dq[i];
}

// BASELINE config 5.  48 + 128 B in, 48 B out, ~10 kflop: fp64-bound.
template <int NJ>
struct OpIkStep
{
    static constexpr int BLOCK = 128, MINB = 1, MAXW = 16;
    ChainConst<NJ> ch;
    const double *q;
    const double *Tt;
    double *q_out;
    double *Vs;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double qi[NJ], m[16], vs[6];
        Rigid target;
        io.load(q, qi);
        io.load(Tt, m);
        rigid_from_record(m, target);
        ik_newton_update<NJ>(ch, target, qi, vs);
        io.store(q_out, qi);
        if (Vs != nullptr)
            io.store(Vs, vs);
    }
};

// The whole loop of test/lie_test.cpp:96-113: the convergence test uses the Vs of the current step
// and the update of that step is still applied, exactly as the reference does.  A warp keeps
// iterating while any of its lanes is unconverged; converged lanes stop updating.
template <int NJ>
struct OpIkSolve
{
    static constexpr int BLOCK = 128, MINB = 1, MAXW = 16;
    ChainConst<NJ> ch;
    const double *q0;
    const double *Tt;
    double *q_out;
    int32_t *iters;
    int8_t *status;
    int max_iter;
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double qi[NJ], m[16], vs[6];
        Rigid target;
        io.load(q0, qi);
        io.load(Tt, m);
        rigid_from_record(m, target);
        bool done = !io.valid;
        int it = 0;
        for (int k = 0; k < max_iter; k++)
        {
            if (__all_sync(0xffffffffu, done))
                break;
            double qn[NJ];
#pragma unroll
            for (int i = 0; i < NJ; i++)
                qn[i] = qi[i];
            ik_newton_update<NJ>(ch, target, qn, vs);
            if (!done)
            {
#pragma unroll
                for (int i = 0; i < NJ; i++)
                    qi[i] = qn[i];
                it++;
                const double ew = sqrt(fma(vs[2], vs[2], fma(vs[1], vs[1], vs[0] * vs[0])));
                const double ev = sqrt(fma(vs[5], vs[5], fma(vs[4], vs[4], vs[3] * vs[3])));
                done = ew < EPS_SP && ev < EPS_SP;
            }
        }
        io.store(q_out, qi);
        if (io.valid)
        {
            if (iters != nullptr)
                iters[io.i] = it;
            if (status != nullptr)
                status[io.i] = done ? PPX_STATUS_CONVERGED : PPX_STATUS_DIVERGED;
        }
    }
};

template <int NJ>
int fk_jacobian_n(const double *screws, const double *home, const double *q, double *T, double *Js, size_t n,
                  int layout, size_t ld, void *stream)
{
    ChainConst<NJ> ch;
    fold_chain<NJ>(screws, home, ch);
    if (T && Js)
        return dispatch(layout, OpFkJacobian<NJ, true, true>{ch, q, T, Js}, n, ld, stream);
    if (T)
        return dispatch(layout, OpFkJacobian<NJ, true, false>{ch, q, T, nullptr}, n, ld, stream);
    return dispatch(layout, OpFkJacobian<NJ, false, true>{ch, q, nullptr, Js}, n, ld, stream);
}

template <int NJ>
int ik_step_n(const double *screws, const double *home, const double *q, const double *Tt, double *q_out,
              double *Vs, size_t n, int layout, size_t ld, void *stream)
{
    ChainConst<NJ> ch;
    fold_chain<NJ>(screws, home, ch);
    return dispatch(layout, OpIkStep<NJ>{ch, q, Tt, q_out, Vs}, n, ld, stream);
}

template <int NJ>
int ik_solve_n(const double *screws, const double *home, const double *q0, const double *Tt, int max_iter,
               double *q_out, int32_t *iters, int8_t *status, size_t n, int layout, size_t ld, void *stream)
{
    ChainConst<NJ> ch;
    fold_chain<NJ>(screws, home, ch);
    return dispatch(layout, OpIkSolve<NJ>{ch, q0, Tt, q_out, iters, status, max_iter}, n, ld, stream);
}

} // namespace

extern "C"
{

int ppx_cuda_batch_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q,
                                   double *T, double *Js, size_t n, int layout, size_t ld, void *stream)
{
    if (njoints < 1 || njoints > PPX_MAX_JOINTS)
        return PPX_ERR_BAD_SHAPE;
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(screws, home, q);
    if (!T && !Js)
        return PPX_ERR_BAD_ARGUMENT;
    switch (njoints)
    {
    case 1: return fk_jacobian_n<1>(screws, home, q, T, Js, n, layout, ld, stream);
    case 2: return fk_jacobian_n<2>(screws, home, q, T, Js, n, layout, ld, stream);
    case 3: return fk_jacobian_n<3>(screws, home, q, T, Js, n, layout, ld, stream);
    case 4: return fk_jacobian_n<4>(screws, home, q, T, Js, n, layout, ld, stream);
    case 5: return fk_jacobian_n<5>(screws, home, q, T, Js, n, layout, ld, stream);
    case 6: return fk_jacobian_n<6>(screws, home, q, T, Js, n, layout, ld, stream);
    case 7: return fk_jacobian_n<7>(screws, home, q, T, Js, n, layout, ld, stream);
    case 8: return fk_jacobian_n<8>(screws, home, q, T, Js, n, layout, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                                  const double *Ttarget, double *q_out, double *Vs, size_t n, int layout,
                                  size_t ld, void *stream)
{
    // linsolve<SVD> on the 6 x N Jacobian returns b.sub<N,1>: N <= 6 (include/linalg.hpp:1208)
    if (njoints < 3 || njoints > 6)
        return PPX_ERR_BAD_SHAPE;
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(screws, home, q, Ttarget, q_out);
    switch (njoints)
    {
    case 3: return ik_step_n<3>(screws, home, q, Ttarget, q_out, Vs, n, layout, ld, stream);
    case 4: return ik_step_n<4>(screws, home, q, Ttarget, q_out, Vs, n, layout, ld, stream);
    case 5: return ik_step_n<5>(screws, home, q, Ttarget, q_out, Vs, n, layout, ld, stream);
    case 6: return ik_step_n<6>(screws, home, q, Ttarget, q_out, Vs, n, layout, ld, stream);
    }
    return PPX_ERR_BAD_SHAPE;
}

int ppx_cuda_batch_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                            const double *Ttarget, int max_iter, double *q_out, int32_t *iters, int8_t *status,
                            size_t n, int layout, size_t ld, void *stream)
{
    if (njoints != 6)
        return PPX_ERR_BAD_SHAPE;
    if (max_iter < 0)
        return PPX_ERR_BAD_ARGUMENT;
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(screws, home, q0, Ttarget, q_out);
    return ik_solve_n<6>(screws, home, q0, Ttarget, max_iter, q_out, iters, status, n, layout, ld, stream);
}

} // extern "C"
