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

// dq = linsolve<SVD>(Js, Vs).x: a square Jacobian (6 joints) takes the row-orthogonalising solve that needs no V
template <int NJ>
__device__ __forceinline__ typename std::enable_if<NJ == 6>::type svd_solve_dispatch(double (&js)[36], const double (&vs)[6],
                                                                                    double (&dq)[6])
{
    jacobi_rowsolve<6>(js, vs, dq);
}
template <int NJ>
__device__ __forceinline__ typename std::enable_if<NJ != 6>::type svd_solve_dispatch(double (&js)[6 * NJ],
                                                                                    const double (&vs)[6], double (&dq)[NJ])
{
    double V[NJ * NJ], w2[NJ];
    jacobi_svd_core<6, NJ, true>(js, V, w2);
    jacobi_svd_solve<6, NJ>(js, V, w2, vs, dq);
}

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
    double dq[NJ];
    svd_solve_dispatch<NJ>(js, vs, dq);
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

// ------------------------------------------------------------------------------------------------
// kinematics<6>::inverseSpace(pose, init) of demo/robotics.hpp:126-152 -- the bounded trust-region
// dogleg details::CoDo<6,6> (include/optimization.hpp:497-829) on f(q) = log(T(q) pose^-1) with the analytic Jacobian
// [[I,0],[-hat(p),I]] Js(q) of robotics.hpp:134-140 and the joints' ranges as the box.  One instance per thread: the
// iteration counts are data dependent, so a warp runs as long as its slowest lane (no warp-level collectives inside).
template <int NJ>
struct CodoProblem
{
    const ChainConst<NJ> &ch;
    Rigid poseI; // pose^-1
    // f and df are deliberately NOT inlined: the optimiser calls f from three places, and one copy of the kinematics
    // keeps the iteration body inside the instruction cache (ncu: `no_instruction` was the top stall when inlined)
    __device__ __noinline__ void f(const double (&q)[NJ], double (&fx)[6]) const
    {
        Rigid t, x;
        poe_fk_jacobian<NJ, true, false>(ch, q, t, [](int, const double(&)[6]) {});
        rigid_mul(t, poseI, x);
        SE3_log(x, fx);
    }
    // column j of [[I,0],[-hat(p),I]] Js = (w_j ; v_j - p x w_j), p = translation of T(q) pose^-1
    __device__ __noinline__ void df(const double (&q)[NJ], double (&jac)[6 * NJ]) const
    {
        Rigid t, x;
        poe_fk_jacobian<NJ, true, true>(ch, q, t, [&](int col, const double(&v)[6]) {
#pragma unroll
            for (int k = 0; k < 6; k++)
                jac[6 * col + k] = v[k];
        });
        rigid_mul(t, poseI, x);
#pragma unroll
        for (int j = 0; j < NJ; j++)
        {
            const double w[3] = {jac[6 * j], jac[6 * j + 1], jac[6 * j + 2]};
            double c[3];
            cross3(x.p, w, c);
#pragma unroll
            for (int k = 0; k < 3; k++)
                jac[6 * j + 3 + k] -= c[k];
        }
    }
};

__device__ __forceinline__ double norm6(const double (&v)[6])
{
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        a = fma(v[i], v[i], a);
    return sqrt(a);
}
__device__ __forceinline__ double dot6(const double (&a)[6], const double (&b)[6])
{
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        r = fma(a[i], b[i], r);
    return r;
}
__device__ __forceinline__ void matvec6(const double (&A)[36], const double (&x)[6], double (&y)[6])
{
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++)
            a = fma(A[i + 6 * k], x[k], a);
        y[i] = a;
    }
}
__device__ __forceinline__ double min6(const double (&v)[6])
{
    double m = v[0];
#pragma unroll
    for (int i = 1; i < 6; i++)
        m = fmin(m, v[i]);
    return m;
}

// CoDo<6,6>::operator() (optimization.hpp:542-779) with DMatrixCi (:787-808): same constants, same tests, same order
// of decisions -- written as a state machine (codo_init + one codo_step per trip of the optimiser's while loop) so that
// a lane whose instance has finished can pick up the next instance instead of idling until the slowest lane of its
// warp is done (trip counts range from 3 to ITMAX = 300; ncu on the plain per-thread loop: 5.4 of 32 lanes active).
struct CodoState
{
    double x[6], fx[6], grad[6], p[6];
    double fnrm, delta0, y;
    unsigned itc;
    int stat;
};

constexpr double CODO_MAX_SP = 3.4028234663852886e+38; // MAX_SP, include/exprtmpl.hpp:14

// -> true when the instance is already finished (illegal start, :546-556)
__device__ __forceinline__ bool codo_init(const CodoProblem<6> &pb, const double (&lo)[6], const double (&hi)[6],
                                          CodoState &st)
{
    bool legal = true;
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        legal = legal && lo[i] <= st.x[i] && st.x[i] <= hi[i];
        st.grad[i] = 0.0;
        st.p[i] = 0.0;
    }
    st.itc = 0;
    st.delta0 = 1.0;
    st.stat = PPX_STATUS_CONVERGED;
    if (!legal)
    {
        st.y = CODO_MAX_SP;
        st.stat = PPX_STATUS_DIVERGED;
        return true;
    }
    pb.f(st.x, st.fx);
    st.fnrm = norm6(st.fx);
    st.y = 0.5 * st.fnrm;
    return false;
}

// one trip of `while (fnrm > FTOLA && itc++ < ITMAX)` (:567-776) -> true when the optimiser is done
__device__ __forceinline__ bool codo_step(const CodoProblem<6> &pb, const double (&lo)[6], const double (&hi)[6],
                                          CodoState &st)
{
    constexpr double MAX_SP = CODO_MAX_SP, sigma = 0.99995, theta = 0.99995, FTOLA = EPS_SP;
    constexpr unsigned ITMAX = 300;
    double(&x)[6] = st.x;
    double(&fx)[6] = st.fx;
    double(&grad)[6] = st.grad;
    double(&p)[6] = st.p;
    double &fnrm = st.fnrm;
    double &delta0 = st.delta0;
    const double sqrt_eps = sqrt(EPS_DP);
    if (!(fnrm > FTOLA && st.itc++ < ITMAX))
    {
        st.y = 0.5 * fnrm;
        return true;
    }
    const unsigned itc = st.itc;
    {
        const double fnrm_old = fnrm;
        double grad_old[6], jac[36], tmp[6], d[6], dsqrt[6], G[6], dgrad[6], pc[6], pn[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
            grad_old[i] = grad[i];
        pb.df(x, jac);
#pragma unroll
        for (int i = 0; i < 6; i++) // grad = jac^T fx
        {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++)
                a = fma(jac[k + 6 * i], fx[k], a);
            grad[i] = a;
        }
        double lambda;
        if (itc == 1)
            lambda = fmax(1e-2, norm6(grad));
        else
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                tmp[i] = grad[i] - grad_old[i];
            lambda = fmax(1e-2, dot6(p, tmp) / dot6(p, p));
        }
#pragma unroll
        for (int i = 0; i < 6; i++) // Hager scaling, :787-808
        {
            if (grad[i] < 0.0 && hi[i] <= MAX_SP)
                d[i] = 1.0 / (lambda - grad[i] / (hi[i] - x[i]));
            else if (grad[i] > 0.0 && lo[i] >= -MAX_SP)
                d[i] = 1.0 / (lambda + grad[i] / (x[i] - lo[i]));
            else
                d[i] = 1.0 / lambda;
        }
        if (min6(d) < EPS_DP)
        {
            st.stat = PPX_STATUS_DIVERGED;
            st.y = 0.5 * fnrm;
            return true;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            dsqrt[i] = sqrt(d[i]);
            G[i] = 1.0 / dsqrt[i];
            dgrad[i] = d[i] * grad[i];
            tmp[i] = G[i] * dgrad[i];
        }
        const double nGdgrad = norm6(tmp);
        if (norm6(dgrad) < FTOLA)
        {
            st.y = 0.5 * fnrm;
            return true;
        }
        double jd[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
            tmp[i] = dsqrt[i] * grad[i];
        matvec6(jac, dgrad, jd);
        const double ratio = norm6(tmp) / norm6(jd);
        const double vert = ratio * ratio;
#pragma unroll
        for (int i = 0; i < 6; i++)
            pc[i] = -vert * dgrad[i];
        if (itc == 1)
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                tmp[i] = grad[i] / G[i];
            delta0 = norm6(tmp);
        }
        // projected Newton step through LU with partial pivoting (linsolve<Factorization::LU>, :608)
        double lu[36], rx[6];
#pragma unroll
        for (int i = 0; i < 36; i++)
            lu[i] = jac[i];
#pragma unroll
        for (int i = 0; i < 6; i++)
            rx[i] = fx[i];
        bool even;
        const bool singular = lu_factor_solve<6, 1>(lu, rx, even);
#pragma unroll
        for (int i = 0; i < 6; i++)
            pn[i] = singular ? -fx[i] : -rx[i];
        if (!singular)
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                pn[i] = fmax(x[i] + pn[i], lo[i]) - x[i];
#pragma unroll
            for (int i = 0; i < 6; i++)
                pn[i] = fmin(x[i] + pn[i], hi[i]) - x[i];
            const double sc = fmax(sigma, 1.0 - norm6(pn));
#pragma unroll
            for (int i = 0; i < 6; i++)
                pn[i] *= sc;
        }
        // trust-region loop, :620-727
        double rho = 0.0, delta1 = delta0, xp[6], fxp[6], fxpnrm = 0.0;
        bool again;
        do
        {
            double pciv[6], alp[6];
            const bool cauchy = vert * nGdgrad > delta1;
#pragma unroll
            for (int i = 0; i < 6; i++)
                pciv[i] = cauchy ? -delta1 / nGdgrad * dgrad[i] : pc[i];
#pragma unroll
            for (int i = 0; i < 6; i++)
                alp[i] = fabs(pciv[i]) < EPS_DP ? MAX_SP : fmax((lo[i] - x[i]) / pciv[i], (hi[i] - x[i]) / pciv[i]);
            double alpha1 = min6(alp);
            if (alpha1 <= 1.0)
            {
                const double sc = fmax(theta, 1.0 - norm6(pciv)) * alpha1;
#pragma unroll
                for (int i = 0; i < 6; i++)
                    pciv[i] *= sc;
            }
            if (!singular)
            {
                double aa[6], seg[6], bb[6], jp[6];
                matvec6(jac, pciv, jp);
#pragma unroll
                for (int i = 0; i < 6; i++)
                {
                    aa[i] = fx[i] + jp[i];
                    seg[i] = pn[i] - pciv[i];
                }
                matvec6(jac, seg, bb);
                double gamma = 0.0;
                if (norm6(bb) != 0.0)
                {
                    const double gamma_min = -dot6(aa, bb) / dot6(bb, bb);
                    double Gseg[6], Gpciv[6];
#pragma unroll
                    for (int i = 0; i < 6; i++)
                    {
                        Gseg[i] = G[i] * seg[i];
                        Gpciv[i] = G[i] * pciv[i];
                    }
                    const double a = dot6(Gseg, Gseg); // SQR(norm2(.))
                    const double b = dot6(Gpciv, Gseg);
                    const double c = dot6(Gpciv, Gpciv) - delta1 * delta1;
                    const double root = sqrt(b * b - a * c);
                    const double l1 = (-b + (-b >= 0.0 ? fabs(root) : -fabs(root))) / a;
                    const double l2 = c / (l1 * a);
                    if (gamma_min > 0.0)
                    {
                        gamma = fmin(gamma_min, fmax(l1, l2));
                        if (gamma > 1.0)
                        {
#pragma unroll
                            for (int i = 0; i < 6; i++)
                                alp[i] = seg[i] != 0.0 ? fmax((lo[i] - x[i] - pciv[i]) / seg[i], (hi[i] - x[i] - pciv[i]) / seg[i])
                                                       : MAX_SP;
                            gamma = fmin(theta * min6(alp), gamma);
                        }
                    }
                    else
                    {
                        gamma = fmax(gamma_min, fmin(l1, l2));
                        if (gamma < 0.0)
                        {
#pragma unroll
                            for (int i = 0; i < 6; i++)
                                alp[i] = seg[i] != 0.0 ? fmax((x[i] + pciv[i] - lo[i]) / seg[i], (x[i] + pciv[i] - hi[i]) / seg[i])
                                                       : MAX_SP;
                            gamma = fmax(-theta * min6(alp), gamma);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; i++)
                    p[i] = (1.0 - gamma) * pciv[i] + gamma * pn[i];
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 6; i++)
                    p[i] = pciv[i];
            }
            double jp2[6], lin[6], Gp[6];
#pragma unroll
            for (int i = 0; i < 6; i++)
                xp[i] = x[i] + p[i];
            pb.f(xp, fxp);
            fxpnrm = norm6(fxp);
            matvec6(jac, p, jp2);
#pragma unroll
            for (int i = 0; i < 6; i++)
            {
                lin[i] = fx[i] + jp2[i];
                Gp[i] = G[i] * p[i];
            }
            rho = (fnrm - fxpnrm) / (fnrm - norm6(lin));
            again = false;
            if (rho < 0.25)
            {
                delta1 = fmin(0.25 * delta0, 0.5 * norm6(Gp));
                again = delta1 > sqrt_eps;
            }
            else if (rho > 0.75)
                delta0 = fmax(delta1, 2.0 * norm6(Gp)); // :735-738 (only reached when the loop ends here)
        } while (again);
        if (rho < 0.25 && delta1 < sqrt_eps)
        {
            st.stat = PPX_STATUS_DIVERGED;
            st.y = 0.5 * fnrm;
            return true;
        }
        bool need_update = false;
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            x[i] = xp[i];
            if (x[i] < lo[i] + 1e3 * EPS_DP)
            {
                x[i] = lo[i] + 1e3 * EPS_DP;
                need_update = true;
            }
            if (x[i] > hi[i] - 1e3 * EPS_DP)
            {
                x[i] = hi[i] - 1e3 * EPS_DP;
                need_update = true;
            }
        }
        if (need_update)
        {
            pb.f(x, fx);
            fnrm = norm6(fx);
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                fx[i] = fxp[i];
            fnrm = fxpnrm;
        }
        st.y = 0.5 * fnrm;
        if (fabs(fnrm - fnrm_old) < FTOLA * fnrm && fnrm > FTOLA)
            return true;
    }
    return false;
}

// Persistent warps with per-lane work refill: every lane owns one instance at a time and, when its optimiser
// finishes, claims the next unprocessed instance from a global counter (one atomicAdd per warp per refill round).
// All lanes of a warp then execute the same codo_step in lockstep, each on its own instance.
struct IkCodoArgs
{
    ChainConst<6> ch;
    double lo[6], hi[6];
    const double *Tt;
    const double *init;
    double *q_out;
    double *x_raw;
    double *fnorm;
    int8_t *status;
};

template <int LAYOUT>
__device__ __forceinline__ double rec_get(const double *base, size_t i, int k, int width, size_t ld)
{
    return LAYOUT == LAYOUT_AOS ? __ldcs(base + i * width + k) : __ldcs(base + (size_t)k * ld + i);
}
template <int LAYOUT>
__device__ __forceinline__ void rec_put(double *base, size_t i, int k, int width, size_t ld, double v)
{
    if (LAYOUT == LAYOUT_AOS)
        __stcs(base + i * width + k, v);
    else
        __stcs(base + (size_t)k * ld + i, v);
}

template <int LAYOUT>
__global__ void __launch_bounds__(128, 2) ik_codo_kernel(const __grid_constant__ IkCodoArgs a, size_t n, size_t ld,
                                                        unsigned long long *next)
{
    const unsigned lane = threadIdx.x & 31;
    CodoProblem<6> pb{a.ch, Rigid()};
    CodoState st;
    double q0[6];
    size_t idx = 0;
    bool busy = false, fresh = false;
    for (;;)
    {
        // refill: lanes without work claim consecutive instances
        const unsigned want = __ballot_sync(0xffffffffu, !busy);
        if (want)
        {
            unsigned long long first = 0;
            if (lane == (unsigned)(__ffs(want) - 1))
                first = atomicAdd(next, (unsigned long long)__popc(want));
            first = __shfl_sync(0xffffffffu, first, __ffs(want) - 1);
            if (!busy)
            {
                idx = (size_t)first + __popc(want & ((1u << lane) - 1));
                if (idx < n)
                {
                    double m[16];
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        m[k] = rec_get<LAYOUT>(a.Tt, idx, k, 16, ld);
#pragma unroll
                    for (int k = 0; k < 6; k++)
                    {
                        q0[k] = rec_get<LAYOUT>(a.init, idx, k, 6, ld);
                        st.x[k] = q0[k];
                    }
                    Rigid pose;
                    rigid_from_record(m, pose);
                    rigid_inv(pose, pb.poseI);
                    busy = true;
                    fresh = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, busy))
            break;
        if (busy)
        {
            bool done = false;
            if (fresh)
            {
                done = codo_init(pb, a.lo, a.hi, st);
                fresh = false;
            }
            if (!done)
                done = codo_step(pb, a.lo, a.hi, st);
            if (done)
            {
#pragma unroll
                for (int k = 0; k < 6; k++)
                {
                    // inverseSpace returns the optimiser's x only if CONVERGED, else init (robotics.hpp:151)
                    rec_put<LAYOUT>(a.q_out, idx, k, 6, ld, st.stat == PPX_STATUS_CONVERGED ? st.x[k] : q0[k]);
                    if (a.x_raw != nullptr)
                        rec_put<LAYOUT>(a.x_raw, idx, k, 6, ld, st.x[k]);
                }
                if (a.fnorm != nullptr)
                    a.fnorm[idx] = st.y;
                if (a.status != nullptr)
                    a.status[idx] = (int8_t)st.stat;
                busy = false;
            }
        }
    }
}
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

int ppx_cuda_batch_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                           const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                           int8_t *status, size_t n, int layout, size_t ld, void *stream)
{
    if (njoints != 6) // CoDo<N,6> solves a square system through linsolve<LU> (optimization.hpp:608)
        return PPX_ERR_BAD_SHAPE;
    if (n == 0)
        return PPX_OK;
    PPX_CHECK_PTRS(screws, home, lo, hi, Ttarget, init, q_out);
    if (layout != PPX_LAYOUT_AOS && layout != PPX_LAYOUT_SOA)
        return PPX_ERR_BAD_LAYOUT;
    if (layout == PPX_LAYOUT_SOA && ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    IkCodoArgs a;
    fold_chain<6>(screws, home, a.ch);
    for (int i = 0; i < 6; i++)
    {
        a.lo[i] = lo[i];
        a.hi[i] = hi[i];
    }
    a.Tt = Ttarget, a.init = init, a.q_out = q_out, a.x_raw = x_raw, a.fnorm = fnorm, a.status = status;
    DeviceInfo info;
    int device = 0;
    int rc = device_info(info, device);
    if (rc != PPX_OK)
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    // the work counter lives for this launch only (stream-ordered allocation, zeroed on the stream)
    unsigned long long *next = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&next, sizeof(*next), s);
    if (e != cudaSuccess)
        return (int)e;
    cudaMemsetAsync(next, 0, sizeof(*next), s);
    auto kern = layout == PPX_LAYOUT_AOS ? ik_codo_kernel<LAYOUT_AOS> : ik_codo_kernel<LAYOUT_SOA>;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0);
    const size_t want = (n + 127) / 128;
    const size_t cap = (size_t)info.sms * (per_sm > 0 ? per_sm : 1);
    kern<<<(unsigned)(want < cap ? want : cap), 128, 0, s>>>(a, n, ld, next);
    g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
    e = cudaPeekAtLastError();
    cudaFreeAsync(next, s);
    return (int)e;
}

} // extern "C"
