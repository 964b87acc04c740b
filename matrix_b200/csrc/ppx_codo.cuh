// ppx_codo.cuh -- kinematics<6>::inverseSpace(pose, init) of demo/robotics.hpp:126-152 for one instance per thread:
// the bounded trust-region dogleg details::CoDo<6,6> (include/optimization.hpp:497-829) on f(q) = log(T(q) pose^-1)
// with the analytic Jacobian [[I,0],[-hat(p),I]] Js(q) of robotics.hpp:134-140 and the joints' ranges as the box.
//
// Shape of the code (what the first version of this kernel got wrong, profiles/r1gh_ncu_inverse_space_codo.txt):
//   * The optimiser's control flow is data dependent (3 .. 300 trips of its while loop, each with an inner
//     trust-region loop), so it is written as a STATE MACHINE around one evaluation: every trip of the device loop
//     evaluates residual and Jacobian at the lane's `probe` point -- the start, a trial point x + p, or the clamped
//     iterate -- and codo_advance() then does whatever the reference does between two evaluations.  All lanes of a
//     warp therefore run the expensive part (6 screw exponentials, the Jacobian, one SE3 log) together, whatever
//     trip each of them is in, and a lane whose instance is finished takes the next one (ik_codo_kernel).
//   * The Jacobian at an accepted trial point is the Jacobian the next trip starts from, so residual and Jacobian
//     come out of ONE pass over the chain (the reference evaluates f and df separately, robotics.hpp:128-140).
//   * The instruction footprint has to stay inside the 32 KB instruction cache of an SM: the first version was
//     130 KB of straight-line code and spent 2/3 of its cycles waiting for instruction fetches.  Hence the chain
//     is a rolled loop over the joints, there is one copy of every block, and divisions / square roots use the
//     branch-free reciprocal (square root) of ppx_math.cuh with a shared out-of-line IEEE fallback for
//     denormal / huge / zero operands.
//
// The file is plain C++ apart from PPX_DEV / __noinline__, so tests/cpp/codo_host.cpp can compile the very same
// state machine for the host and replay the reference's fixtures through it without a GPU.
#pragma once
#include <type_traits>

#include "../../include/ppx_cuda.h"
#include "ppx_math.cuh"

namespace ppx_dev
{

constexpr double CODO_MAX_SP = 3.4028234663852886e+38; // MAX_SP, include/exprtmpl.hpp:14

// ---- compact arithmetic ------------------------------------------------------------------------------------------
// Divisions and square roots come in two flavours.  Where the operand is known to be an ordinary number (or the
// result is discarded by a select when it is not) the branch-free reciprocal / reciprocal square root of ppx_math.cuh
// is used as is: ~8 instructions, no call.  Where the reference's behaviour on 0, denormal, huge or negative operands
// matters (rho = 0/0 at a stalled iterate, the discriminant of the dogleg quadratic) the operand's exponent is tested
// and the rare case goes to ONE out-of-line IEEE routine.
PPX_DEV int codo_hiword(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    long long u;
    __builtin_memcpy(&u, &x, 8);
    return (int)(u >> 32);
#endif
}
static __device__ __noinline__ double codo_div_ieee(double a, double b) { return a / b; }
static __device__ __noinline__ double codo_sqrt_ieee(double a) { return sqrt(a); }

// a / b for b != 0 of ordinary magnitude (2^-960 < |b| < 2^960); anything else gives NaN or inf, never a trap
PPX_DEV double codo_div_fast(double a, double b) { return a * rcp_nr(b); }
// a / b with IEEE results for every b
PPX_DEV double codo_div(double a, double b)
{
    const bool mid = (unsigned)(((codo_hiword(b) >> 20) & 0x7ff) - 64) < 1920u; // biased exponent in [64, 1983]
    return mid ? a * rcp_nr(b) : codo_div_ieee(a, b);
}
// sqrt(a) for a >= 0 (sums of squares); below 2^-960 the result is 0, which every caller treats like the true 1e-145
PPX_DEV double codo_sqrt_fast(double a) { return codo_hiword(a) >= (64 << 20) ? a * rsqrt_nr(a) : 0.0; }
// sqrt(a) with IEEE results for every a (NaN for a < 0)
PPX_DEV double codo_sqrt(double a)
{
    const int hi = codo_hiword(a);
    return (hi >= (64 << 20) && hi < (1984 << 20)) ? a * rsqrt_nr(a) : codo_sqrt_ieee(a);
}

// Per-lane arrays that are not worth registers live in shared memory in the kernel, interleaved by thread: element k of
// this lane is p[k * CODO_JS] with CODO_JS = the CTA size (conflict-free; and unlike the stack it cannot be evicted to
// L2 / DRAM: the first version of this kernel moved 10x its algorithmic bytes through DRAM as local-memory traffic).
// These are the two Jacobians (indexed by the running joint of the rolled chain loop) and the vectors the optimiser
// carries from one evaluation to the next but does not touch during it.  The host build uses CODO_JS = 1.
#ifndef PPX_CODO_JSTRIDE
#define PPX_CODO_JSTRIDE 1
#endif
constexpr int CODO_JS = PPX_CODO_JSTRIDE;
struct LaneArray
{
    double *p;
    PPX_DEV double &operator[](int k) const { return p[(size_t)k * CODO_JS]; }
    PPX_DEV LaneArray from(int k) const { return LaneArray{p + (size_t)k * CODO_JS}; }
};

template <class A, class B>
PPX_DEV double codo_dot6(const A &a, const B &b)
{
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
        r = fma(a[i], b[i], r);
    return r;
}
template <class A>
PPX_DEV double codo_norm6(const A &v)
{
    return codo_sqrt_fast(codo_dot6(v, v));
}
template <class A>
PPX_DEV double codo_min6(const A &v)
{
    double m = v[0];
#pragma unroll
    for (int i = 1; i < 6; i++)
        m = fmin(m, v[i]);
    return m;
}
// y = J x, J column-major 6 x 6
template <class X>
PPX_DEV void codo_matvec(const LaneArray J, const X &x, double (&y)[6])
{
#pragma unroll
    for (int i = 0; i < 6; i++)
        y[i] = J[i] * x[0];
#pragma unroll
    for (int k = 1; k < 6; k++)
#pragma unroll
        for (int i = 0; i < 6; i++)
            y[i] = fma(J[i + 6 * k], x[k], y[i]);
}

// ---- residual and Jacobian in one rolled pass over the chain -----------------------------------------------------
// fx = log(T(q) pose^-1) (robotics.hpp:128-132) and jac = [[I,0],[-hat(p),I]] Js(q), p = translation of
// T(q) pose^-1 (robotics.hpp:134-140): column j is (w_j ; v_j - p x w_j) with (w_j ; v_j) = Ad(P_j) S_j.
// The prefix P starts as the identity, for which Ad(P) S and P E are exact, so the rolled loop produces the bits of
// the unrolled poe_fk_jacobian.
// With a FixedKinds pattern (ppx_math.cuh) the loop is unrolled instead and every joint compiles to the variant its
// axis allows: ~330 instead of ~600 fp64 instructions and no constant-bank indexing, in about the same code size.
template <int NJ, class Kinds = RuntimeKinds>
PPX_DEV void codo_eval(const ChainConst<NJ> &ch, const Rigid &poseI, const double *q, double (&fx)[6], const LaneArray jac)
{
    Rigid P;
    if (std::is_same<Kinds, RuntimeKinds>::value)
    {
        rigid_identity(P);
#pragma unroll 1
        for (int i = 0; i < NJ; i++)
        {
            const JointConst &jc = ch.joint[i];
            double col[6];
            rigid_adjoint_apply(P, jc.w, jc.v, col);
#pragma unroll
            for (int k = 0; k < 6; k++)
                jac[6 * i + k] = col[k];
            Rigid E, Q;
            screw_exp(jc, q[i], E);
            rigid_mul(P, E, Q);
            P = Q;
        }
    }
    else
    {
        auto emit = [&](int col, const double(&v)[6]) {
#pragma unroll
            for (int k = 0; k < 6; k++)
                jac[6 * col + k] = v[k];
        };
#pragma unroll
        for (int i = 0; i < NJ; i++)
        {
            switch (Kinds::get(ch, i))
            {
            case 1: poe_joint<1>(ch.joint[i], q[i], P, i, i == 0, true, true, emit); break;
            case 2: poe_joint<2>(ch.joint[i], q[i], P, i, i == 0, true, true, emit); break;
            case 3: poe_joint<3>(ch.joint[i], q[i], P, i, i == 0, true, true, emit); break;
            default: poe_joint<0>(ch.joint[i], q[i], P, i, i == 0, true, true, emit); break;
            }
        }
    }
    Rigid M, T, X;
#pragma unroll
    for (int i = 0; i < 9; i++)
        M.r[i] = ch.home_r[i];
#pragma unroll
    for (int i = 0; i < 3; i++)
        M.p[i] = ch.home_p[i];
    rigid_mul(P, M, T);
    rigid_mul(T, poseI, X);
#pragma unroll 1
    for (int j = 0; j < NJ; j++)
    {
        const double w[3] = {jac[6 * j], jac[6 * j + 1], jac[6 * j + 2]};
        double c[3];
        cross3(X.p, w, c);
#pragma unroll
        for (int k = 0; k < 3; k++)
            jac[6 * j + 3 + k] -= c[k];
    }
    SE3_log(X, fx);
}

// ---- CoDo<6,6>::operator() (optimization.hpp:542-779) with DMatrixCi (:787-808) as a state machine -----------------
enum CodoPhase
{
    CODO_START = 0, // probe = the start point: the evaluation gives fx, jac of the first trip (:558-560)
    CODO_TRIAL = 1, // probe = x + p of the trust-region loop (:620-727)
    CODO_CLAMPED = 2 // probe = the accepted point after it was pushed back inside the box (:747-762)
};

struct CodoLane
{
    double x[6], fx[6];  // iterate and residual there
    LaneArray jbuf;      // 2 x 36 doubles owned by the caller: the Jacobian at x is jbuf[36 cur ..], the evaluation at
    int cur;             // `probe` writes the other half, and accepting a trial point is a flip of `cur`
    double probe[6];             // where the next evaluation happens
    LaneArray grad, p;           // carried from trip to trip (Barzilai-Borwein lambda, :574-583)
    LaneArray pn, dgrad, G;      // of the running trip; needed again every time the trust region shrinks
    double fnrm, fnrm_old, delta0, delta1, vert, nGdgrad, y;
    unsigned itc;
    int stat, phase;
    bool singular;
};

// hands the lane its CODO_LANE_DOUBLES doubles of interleaved storage
constexpr int CODO_LANE_DOUBLES = 72 + 5 * 6;
PPX_DEV void codo_attach(CodoLane &s, LaneArray mem)
{
    s.jbuf = mem;
    s.grad = mem.from(72), s.p = mem.from(78), s.pn = mem.from(84), s.dgrad = mem.from(90), s.G = mem.from(96);
}
PPX_DEV LaneArray codo_jac(const CodoLane &s) { return s.jbuf.from(36 * s.cur); }
PPX_DEV LaneArray codo_jac_probe(const CodoLane &s) { return s.jbuf.from(36 * (s.cur ^ 1)); }

// after x was set: false -> evaluate at probe; true -> finished (illegal start, :546-556)
PPX_DEV bool codo_begin(const double (&lo)[6], const double (&hi)[6], CodoLane &s)
{
    bool legal = true;
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        legal = legal && lo[i] <= s.x[i] && s.x[i] <= hi[i];
        s.grad[i] = 0.0;
        s.p[i] = 0.0;
        s.probe[i] = s.x[i];
    }
    s.itc = 0;
    s.cur = 0;
    s.delta0 = 1.0;
    s.stat = PPX_STATUS_CONVERGED;
    s.phase = CODO_START;
    if (!legal)
    {
        s.y = CODO_MAX_SP;
        s.stat = PPX_STATUS_DIVERGED;
        return true;
    }
    return false;
}

// The optimiser between two evaluations, in three stages.  Which stages a lane runs depends on its phase and on the
// outcome of the trial, so the lanes of a warp part ways; codo_advance() brings them back together before every stage
// (without that, lanes that reach a stage over different paths would execute it one group after the other).
enum CodoNext
{
    CODO_EVAL = 0, // nothing more to do before the next evaluation
    CODO_DONE = 1, // the optimiser has returned: s.x, s.y, s.stat are its result
    CODO_HEAD = 2, // start the next trip
    CODO_BUILD = 4 // build the dogleg step for the current trust-region radius
};

// stage 1: fp and codo_jac_probe(s) = residual and Jacobian at s.probe have arrived
PPX_DEV unsigned codo_judge(const double (&lo)[6], const double (&hi)[6], CodoLane &s, const double (&fp)[6])
{
    constexpr double FTOLA = EPS_SP;
    constexpr double sqrt_eps = 1.4901161193847656e-08; // sqrt(EPS_DP), exact
    double(&x)[6] = s.x;
    double(&fx)[6] = s.fx;
    double fnew;
    if (s.phase == CODO_TRIAL)
    {
        // judge the trial point, :705-727
        const double fxpnrm = codo_norm6(fp);
        double jp2[6], lin[6], Gp[6];
        codo_matvec(codo_jac(s), s.p, jp2);
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            lin[i] = fx[i] + jp2[i];
            Gp[i] = s.G[i] * s.p[i];
        }
        const double rho = codo_div(s.fnrm - fxpnrm, s.fnrm - codo_norm6(lin));
        if (rho < 0.25)
        {
            s.delta1 = fmin(0.25 * s.delta0, 0.5 * codo_norm6(Gp));
            if (s.delta1 > sqrt_eps)
                return CODO_BUILD; // same trip, smaller trust region
            if (s.delta1 < sqrt_eps)
            {
                s.stat = PPX_STATUS_DIVERGED;
                s.y = 0.5 * s.fnrm;
                return CODO_DONE;
            }
        }
        else if (rho > 0.75)
            s.delta0 = fmax(s.delta1, 2.0 * codo_norm6(Gp)); // :735-738
        // accept, then keep the iterate strictly inside the box (:747-762)
        bool need_update = false;
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            double xi = s.probe[i];
            if (xi < lo[i] + 1e3 * EPS_DP)
            {
                xi = lo[i] + 1e3 * EPS_DP;
                need_update = true;
            }
            if (xi > hi[i] - 1e3 * EPS_DP)
            {
                xi = hi[i] - 1e3 * EPS_DP;
                need_update = true;
            }
            x[i] = xi;
            s.probe[i] = xi;
        }
        if (need_update)
        {
            s.phase = CODO_CLAMPED; // the reference evaluates f again at the clamped point
            return CODO_EVAL;
        }
        fnew = fxpnrm;
    }
    else
        fnew = codo_norm6(fp);
    // residual and Jacobian at the probe belong to the iterate now
#pragma unroll
    for (int i = 0; i < 6; i++)
        fx[i] = fp[i];
    s.cur ^= 1;
    s.fnrm = fnew;
    s.y = 0.5 * fnew;
    // end of a trip, :764-775 (not after the very first evaluation)
    if (s.phase != CODO_START && fabs(s.fnrm - s.fnrm_old) < FTOLA * s.fnrm && s.fnrm > FTOLA)
        return CODO_DONE;
    return CODO_HEAD;
}

// stage 2: `while (fnrm > FTOLA && itc++ < ITMAX)` and the trip up to its trust-region loop, :567-618
PPX_DEV unsigned codo_trip_head(const double (&lo)[6], const double (&hi)[6], CodoLane &s)
{
    constexpr double MAX_SP = CODO_MAX_SP, sigma = 0.99995, FTOLA = EPS_SP;
    constexpr unsigned ITMAX = 300;
    double(&x)[6] = s.x;
    double(&fx)[6] = s.fx;
    const LaneArray jac = codo_jac(s);
    {
        if (!(s.fnrm > FTOLA && s.itc++ < ITMAX))
        {
            s.y = 0.5 * s.fnrm;
            return CODO_DONE;
        }
        s.fnrm_old = s.fnrm;
        double grad_new[6], tmp[6], d[6], dsqrt[6];
#pragma unroll
        for (int i = 0; i < 6; i++) // grad = jac^T fx
        {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++)
                a = fma(jac[k + 6 * i], fx[k], a);
            grad_new[i] = a;
        }
        double lambda;
        if (s.itc == 1)
            lambda = fmax(1e-2, codo_norm6(grad_new));
        else
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                tmp[i] = grad_new[i] - s.grad[i];
            lambda = fmax(1e-2, codo_div(codo_dot6(s.p, tmp), codo_dot6(s.p, s.p)));
        }
#pragma unroll
        for (int i = 0; i < 6; i++) // Hager scaling, :787-808
        {
            const double g = grad_new[i];
            s.grad[i] = g;
            const bool up = g < 0.0 && hi[i] <= MAX_SP, dn = g > 0.0 && lo[i] >= -MAX_SP;
            // lambda - g/(hi-x)  |  lambda + g/(x-lo)  |  lambda
            const double num = up ? -g : g, den = up ? hi[i] - x[i] : x[i] - lo[i];
            // on the bound itself (a legal start) the reference divides by zero: d = 1 / inf = 0 -> DIVERGED below
            const bool on_bound = (up || dn) && !(den >= 1e-290);
            const double bend = (up || dn) ? num * rcp_nr(den) : 0.0;
            // lambda >= 1e-2.  A denominator that overflowed (huge gradient next to a bound) gives the reference
            // d = 1 / inf = 0 and the DIVERGED exit below; the branch-free reciprocal would return NaN there, which
            // fmin() drops -- so a non-finite denominator is mapped to d = 0 explicitly (NaN inputs end the same way)
            const double dd = lambda + bend;
            d[i] = (on_bound || !(fabs(dd) <= DBL_MAX)) ? 0.0 : rcp_nr(dd);
        }
        if (codo_min6(d) < EPS_DP)
        {
            s.stat = PPX_STATUS_DIVERGED;
            s.y = 0.5 * s.fnrm;
            return CODO_DONE;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            s.G[i] = rsqrt_nr(d[i]); // d >= EPS_DP here.  G = 1 / sqrt(d)
            dsqrt[i] = d[i] * s.G[i];
            s.dgrad[i] = d[i] * s.grad[i];
            tmp[i] = s.G[i] * s.dgrad[i];
        }
        s.nGdgrad = codo_norm6(tmp);
        if (codo_norm6(s.dgrad) < FTOLA)
        {
            s.y = 0.5 * s.fnrm;
            return CODO_DONE;
        }
        double jd[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
            tmp[i] = dsqrt[i] * s.grad[i];
        codo_matvec(jac, s.dgrad, jd);
        const double ratio = codo_div(codo_norm6(tmp), codo_norm6(jd));
        s.vert = ratio * ratio;
        if (s.itc == 1)
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                tmp[i] = s.grad[i] * dsqrt[i]; // grad / G
            s.delta0 = codo_norm6(tmp);
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
        s.singular = lu_factor_solve<6, 1>(lu, rx, even);
#pragma unroll
        for (int i = 0; i < 6; i++)
            s.pn[i] = s.singular ? -fx[i] : -rx[i];
        if (!s.singular)
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.pn[i] = fmax(x[i] + s.pn[i], lo[i]) - x[i];
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.pn[i] = fmin(x[i] + s.pn[i], hi[i]) - x[i];
            const double sc = fmax(sigma, 1.0 - codo_norm6(s.pn));
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.pn[i] *= sc;
        }
        s.delta1 = s.delta0;
    }
    return CODO_BUILD;
}

// stage 3: the dogleg step inside the trust region of radius delta1, :622-703; its end point is the next probe
PPX_DEV void codo_build(const double (&lo)[6], const double (&hi)[6], CodoLane &s)
{
    constexpr double MAX_SP = CODO_MAX_SP, theta = 0.99995;
    double(&x)[6] = s.x;
    double(&fx)[6] = s.fx;
    const LaneArray jac = codo_jac(s);
    {
        double pciv[6], alp[6];
        const bool cauchy = s.vert * s.nGdgrad > s.delta1;
        const double shrink = -codo_div(s.delta1, s.nGdgrad);
#pragma unroll
        for (int i = 0; i < 6; i++)
            pciv[i] = cauchy ? shrink * s.dgrad[i] : -s.vert * s.dgrad[i]; // else the Cauchy point pc
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            const double r = rcp_nr(pciv[i]); // NaN / inf for a zero component: discarded by the select
            alp[i] = fabs(pciv[i]) < EPS_DP ? MAX_SP : fmax((lo[i] - x[i]) * r, (hi[i] - x[i]) * r);
        }
        const double alpha1 = codo_min6(alp);
        if (alpha1 <= 1.0)
        {
            const double sc = fmax(theta, 1.0 - codo_norm6(pciv)) * alpha1;
#pragma unroll
            for (int i = 0; i < 6; i++)
                pciv[i] *= sc;
        }
        if (!s.singular)
        {
            double aa[6], seg[6], bb[6], jpc[6];
            codo_matvec(jac, pciv, jpc);
#pragma unroll
            for (int i = 0; i < 6; i++)
            {
                aa[i] = fx[i] + jpc[i];
                seg[i] = s.pn[i] - pciv[i];
            }
            codo_matvec(jac, seg, bb);
            double gamma = 0.0;
            const double bb2 = codo_dot6(bb, bb);
            if (bb2 != 0.0) // norm2(bb) != 0
            {
                const double gamma_min = -codo_div(codo_dot6(aa, bb), bb2);
                double Gseg[6], Gpciv[6];
#pragma unroll
                for (int i = 0; i < 6; i++)
                {
                    Gseg[i] = s.G[i] * seg[i];
                    Gpciv[i] = s.G[i] * pciv[i];
                }
                const double a = codo_dot6(Gseg, Gseg); // SQR(norm2(.))
                const double b = codo_dot6(Gpciv, Gseg);
                const double c = codo_dot6(Gpciv, Gpciv) - s.delta1 * s.delta1;
                const double root = codo_sqrt(b * b - a * c);
                const double l1 = codo_div(-b + (-b >= 0.0 ? fabs(root) : -fabs(root)), a);
                const double l2 = codo_div(c, l1 * a);
                const bool fwd = gamma_min > 0.0;
                gamma = fwd ? fmin(gamma_min, fmax(l1, l2)) : fmax(gamma_min, fmin(l1, l2));
                if (fwd ? gamma > 1.0 : gamma < 0.0)
                {
                    // stay inside the box along the segment, :668-676 / :683-691
#pragma unroll
                    for (int i = 0; i < 6; i++)
                    {
                        const double r = rcp_nr(seg[i]);
                        const double tl = fwd ? lo[i] - x[i] - pciv[i] : x[i] + pciv[i] - lo[i];
                        const double th = fwd ? hi[i] - x[i] - pciv[i] : x[i] + pciv[i] - hi[i];
                        alp[i] = seg[i] != 0.0 ? fmax(tl * r, th * r) : MAX_SP;
                    }
                    const double lim = theta * codo_min6(alp);
                    gamma = fwd ? fmin(lim, gamma) : fmax(-lim, gamma);
                }
            }
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.p[i] = (1.0 - gamma) * pciv[i] + gamma * s.pn[i];
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 6; i++)
                s.p[i] = pciv[i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
            s.probe[i] = x[i] + s.p[i];
        s.phase = CODO_TRIAL;
    }
}

#ifdef __CUDA_ARCH__
#define PPX_CODO_CONVERGE(mask) __syncwarp(mask)
#else
#define PPX_CODO_CONVERGE(mask) ((void)(mask))
#endif

// fp and codo_jac_probe(s) = residual and Jacobian at s.probe.  Does what the reference does up to its next evaluation
// and leaves the point of that evaluation in s.probe; -> true when the optimiser has returned (s.x, s.y, s.stat are its
// result).  `lanes` = the lanes of the warp that make this call together.
PPX_DEV bool codo_advance(unsigned lanes, const double (&lo)[6], const double (&hi)[6], CodoLane &s,
                          const double (&fp)[6])
{
    unsigned next = codo_judge(lo, hi, s, fp);
    PPX_CODO_CONVERGE(lanes);
    if (next & CODO_HEAD)
        next = codo_trip_head(lo, hi, s);
    PPX_CODO_CONVERGE(lanes);
    if (next & CODO_BUILD)
        codo_build(lo, hi, s);
    return (next & CODO_DONE) != 0;
}

} // namespace ppx_dev
