// ppx_math.cuh -- per-instance fp64 device math for the ppx hot path (sm_100a).
//
// One instance lives entirely in the registers of one thread: every array below has
// compile-time extents and every loop is fully unrolled, so nothing is indexed dynamically
// and nothing spills to local memory.  Semantics follow the reference routine cited beside
// each function (file:line under the reference tree); the arithmetic is re-derived for the
// GPU (closed forms instead of 4x4 matrix powers, reciprocals shared between coefficients,
// rotation-aware products) and agrees with the reference to fp64 rounding, which the parity
// tests bound at 1e-12 relative.  Branch thresholds are the reference's, exactly.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstring>

namespace ppx_dev
{

constexpr double EPS_SP = 1.1920928955078125e-07; // float epsilon, include/exprtmpl.hpp:12
constexpr double EPS_DP = 2.220446049250313e-16;  // double epsilon, include/exprtmpl.hpp:13
constexpr double PI = 3.141592653589793;          // include/exprtmpl.hpp:11

#define PPX_DEV __device__ __forceinline__

// Branch-free reciprocal square root / reciprocal: the hardware seed (MUFU.RSQ64H / MUFU.RCP64H, relative
// error ~2^-22) refined by one third-order step, full fp64 accuracy for normal, non-zero arguments.  Unlike
// sqrt() and operator/ there is no special-case slow path, hence no call and no convergence barrier in the
// instruction stream: the independent rotation chains of a Jacobi round stay in one basic block and overlap.
// (x = 0 gives inf/NaN; callers discard that lane's result with a select.)
#ifndef __CUDA_ARCH__
// host build of the device math (tests/cpp): a seed of the hardware's quality -- the exact value with its mantissa cut
// to 23 bits -- over the whole fp64 exponent range (a cast to float would overflow below 1e-77)
inline double ppx_host_seed(double y)
{
    unsigned long long u;
    memcpy(&u, &y, 8);
    u &= 0xFFFFFFFFE0000000ull;
    memcpy(&y, &u, 8);
    return y;
}
#endif
PPX_DEV double rsqrt_nr(double x)
{
    double y;
#ifdef __CUDA_ARCH__
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = ppx_host_seed(1.0 / sqrt(x));
#endif
    const double e = fma(-(x * y), y, 1.0); // 1 - x y^2
    return fma(y * e, fma(0.375, e, 0.5), y); // y (1 + e/2 + 3 e^2/8)
}
PPX_DEV double rcp_nr(double x)
{
    double y;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = ppx_host_seed(1.0 / x);
#endif
    const double e = fma(-x, y, 1.0); // 1 - x y
    return fma(y, fma(e, e, e), y);   // y (1 + e + e^2)
}

// fp32 counterparts for the single-precision preconditioning sweeps of the Jacobi SVD (full fp32 accuracy is plenty:
// that phase only has to find an approximately diagonalising basis)
PPX_DEV float rsqrt_nr(float x)
{
#ifdef __CUDA_ARCH__
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); // one MUFU, 2 ulp
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}
PPX_DEV float rcp_nr(float x)
{
#ifdef __CUDA_ARCH__
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); // one MUFU, 1 ulp
    return y;
#else
    return 1.0f / x;
#endif
}
PPX_DEV float t_fma(float a, float b, float c) { return fmaf(a, b, c); }
PPX_DEV double t_fma(double a, double b, double c) { return fma(a, b, c); }
PPX_DEV float t_abs(float a) { return fabsf(a); }
PPX_DEV double t_abs(double a) { return fabs(a); }

// Rigid transform: rotation r (column-major, r[i + 3 j]) and translation p.  An SE3 record is
// [r0 r1 r2 0 | r3 r4 r5 0 | r6 r7 r8 0 | p0 p1 p2 1] (include/liegroup.hpp:186-200).
struct Rigid
{
    double r[9];
    double p[3];
};

PPX_DEV void rigid_identity(Rigid &t)
{
#pragma unroll
    for (int i = 0; i < 9; i++)
        t.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
    t.p[0] = t.p[1] = t.p[2] = 0.0;
}

PPX_DEV void rigid_from_record(const double (&m)[16], Rigid &t)
{
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++)
            t.r[r + 3 * c] = m[r + 4 * c];
    t.p[0] = m[12], t.p[1] = m[13], t.p[2] = m[14];
}

PPX_DEV void rigid_to_record(const Rigid &t, double (&m)[16])
{
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
#pragma unroll
        for (int r = 0; r < 3; r++)
            m[r + 4 * c] = t.r[r + 3 * c];
        m[3 + 4 * c] = 0.0;
    }
    m[12] = t.p[0], m[13] = t.p[1], m[14] = t.p[2], m[15] = 1.0;
}

// y = R x
PPX_DEV void rot_apply(const double (&r)[9], const double (&x)[3], double (&y)[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
        y[i] = fma(r[i + 6], x[2], fma(r[i + 3], x[1], r[i] * x[0]));
}

// y = R^T x
PPX_DEV void rot_apply_t(const double (&r)[9], const double (&x)[3], double (&y)[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
        y[i] = fma(r[3 * i + 2], x[2], fma(r[3 * i + 1], x[1], r[3 * i] * x[0]));
}

PPX_DEV void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3])
{
    c[0] = fma(a[1], b[2], -a[2] * b[1]);
    c[1] = fma(a[2], b[0], -a[0] * b[2]);
    c[2] = fma(a[0], b[1], -a[1] * b[0]);
}

// C = A * B for rigid transforms (SE3::operator*, include/liegroup.hpp:204-207); the constant
// bottom row of the 4x4 product is not recomputed: 27 + 9 multiply-adds instead of 64.
PPX_DEV void rigid_mul(const Rigid &a, const Rigid &b, Rigid &c)
{
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            c.r[i + 3 * j] = fma(a.r[i + 6], b.r[2 + 3 * j], fma(a.r[i + 3], b.r[1 + 3 * j], a.r[i] * b.r[3 * j]));
#pragma unroll
    for (int i = 0; i < 3; i++)
        c.p[i] = fma(a.r[i + 6], b.p[2], fma(a.r[i + 3], b.p[1], fma(a.r[i], b.p[0], a.p[i])));
}

// SE3::I, include/liegroup.hpp:227-230: (R^T, -R^T p)
PPX_DEV void rigid_inv(const Rigid &a, Rigid &b)
{
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            b.r[i + 3 * j] = a.r[j + 3 * i];
    double t[3];
    rot_apply_t(a.r, a.p, t);
    b.p[0] = -t[0], b.p[1] = -t[1], b.p[2] = -t[2];
}

// SE3::Adt, include/liegroup.hpp:217-226, as a dense 6x6 record (column-major, ad[r + 6 c]):
// [[R, 0], [hat(p) R, R]].
PPX_DEV void rigid_adjoint_record(const Rigid &t, double (&ad)[36])
{
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        const double col[3] = {t.r[3 * c], t.r[3 * c + 1], t.r[3 * c + 2]};
        double pc[3];
        cross3(t.p, col, pc);
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
            ad[r + 6 * c] = col[r];
            ad[r + 3 + 6 * c] = pc[r];
            ad[r + 6 * (c + 3)] = 0.0;
            ad[r + 3 + 6 * (c + 3)] = col[r];
        }
    }
}

// Ad(T) * (w; v) without forming the 6x6: (R w ; p x (R w) + R v)
PPX_DEV void rigid_adjoint_apply(const Rigid &t, const double (&w)[3], const double (&v)[3], double (&out)[6])
{
    double rw[3], rv[3], c[3];
    rot_apply(t.r, w, rw);
    rot_apply(t.r, v, rv);
    cross3(t.p, rw, c);
    out[0] = rw[0], out[1] = rw[1], out[2] = rw[2];
    out[3] = c[0] + rv[0], out[4] = c[1] + rv[1], out[5] = c[2] + rv[2];
}

// so3::exp, include/liegroup.hpp:116-133 (identity below near_zero = 1e-6, include/matrixs.hpp:39-42).
PPX_DEV void so3_exp(const double (&w)[3], double (&r)[9])
{
    const double t2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0]));
    // theta = t2 / sqrt(t2) and 1/theta^2 from ONE branch-free reciprocal square root (no sqrt(), no division: both
    // carry slow-path branches); below the reference's threshold its value is not used
    const bool tiny = t2 < 1.0e-12; // |theta| < 1e-6
    const double rth = rsqrt_nr(tiny ? 1.0 : t2);
    const double theta = tiny ? 0.0 : t2 * rth;
    double s, c;
    sincos(theta, &s, &c);
    const double inv = tiny ? 0.0 : rth * rth;
    const double a = tiny ? 0.0 : s * rth; // sin(theta)/theta
    const double b = (1.0 - c) * inv;              // (1-cos(theta))/theta^2
    // I + a W + b W^2,  W^2 = w w^T - theta^2 I,  1 - b theta^2 = cos(theta)
    const double d = tiny ? 1.0 : c;
    const double bw0 = b * w[0], bw1 = b * w[1], bw2 = b * w[2];
    r[0] = fma(bw0, w[0], d);
    r[4] = fma(bw1, w[1], d);
    r[8] = fma(bw2, w[2], d);
    r[1] = fma(bw0, w[1], a * w[2]);
    r[3] = fma(bw0, w[1], -a * w[2]);
    r[2] = fma(bw0, w[2], -a * w[1]);
    r[6] = fma(bw0, w[2], a * w[1]);
    r[5] = fma(bw1, w[2], a * w[0]);
    r[7] = fma(bw1, w[2], -a * w[0]);
}

// SO3::log, include/liegroup.hpp:60-92, all three branches.
PPX_DEV void SO3_log(const double (&r)[9], double (&w)[3])
{
    const double acosinput = ((r[0] + r[4]) + r[8] - 1.0) * 0.5;
    if (acosinput >= 1.0)
    {
        w[0] = w[1] = w[2] = 0.0;
    }
    else if (acosinput <= -1.0)
    {
        double s, v0, v1, v2;
        if (!(fabs(1.0 + r[8]) < EPS_SP))
        {
            s = 1.0 / sqrt(2.0 + 2.0 * r[8]);
            v0 = r[6], v1 = r[7], v2 = 1.0 + r[8];
        }
        else if (!(fabs(1.0 + r[4]) < EPS_SP))
        {
            s = 1.0 / sqrt(2.0 + 2.0 * r[4]);
            v0 = r[3], v1 = 1.0 + r[4], v2 = r[5];
        }
        else
        {
            s = 1.0 / sqrt(2.0 + 2.0 * r[0]);
            v0 = 1.0 + r[0], v1 = r[1], v2 = r[2];
        }
        w[0] = PI * (s * v0), w[1] = PI * (s * v1), w[2] = PI * (s * v2);
    }
    else
    {
        const double theta = acos(acosinput);
        // sin(acos(c)) = sqrt((1-c)(1+c)): one sqrt instead of a second transcendental
        const double k = 0.5 * theta * rsqrt_nr((1.0 - acosinput) * (1.0 + acosinput)); // theta / (2 sin theta)
        w[0] = (r[5] - r[7]) * k;
        w[1] = (r[6] - r[2]) * k;
        w[2] = (r[1] - r[3]) * k;
    }
}

// se3::exp, include/liegroup.hpp:253-269.  Closed form of I + X + c1 X^2 + c2 X^3 for X = hat(xi):
//   R = cos(phi) I + (sin(phi)/phi) W + c1 w w^T,   p = v + c1 (w x v) + c2 (w x (w x v)),
// with c1 = (1-cos phi)/phi^2, c2 = (phi - sin phi)/phi^3; R = I, p = v when |phi| < EPS_SP (:258-261).
PPX_DEV void se3_exp(const double (&xi)[6], Rigid &t)
{
    const double w[3] = {xi[0], xi[1], xi[2]};
    const double v[3] = {xi[3], xi[4], xi[5]};
    const double p2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0]));
    const bool tiny = p2 < EPS_SP * EPS_SP; // |phi| < EPS_SP
    const double iphi = tiny ? 0.0 : rsqrt_nr(tiny ? 1.0 : p2); // 1/phi: phi and 1/phi^2 follow without sqrt() or division
    const double phi = p2 * iphi;
    double s, c;
    sincos(phi, &s, &c);
    const double inv = iphi * iphi;
    const double a = tiny ? 0.0 : s * iphi;
    const double c1 = (1.0 - c) * inv;
    const double c2 = (phi - s) * inv * iphi;
    const double d = tiny ? 1.0 : c;
    const double bw0 = c1 * w[0], bw1 = c1 * w[1], bw2 = c1 * w[2];
    t.r[0] = fma(bw0, w[0], d);
    t.r[4] = fma(bw1, w[1], d);
    t.r[8] = fma(bw2, w[2], d);
    t.r[1] = fma(bw0, w[1], a * w[2]);
    t.r[3] = fma(bw0, w[1], -a * w[2]);
    t.r[2] = fma(bw0, w[2], -a * w[1]);
    t.r[6] = fma(bw0, w[2], a * w[1]);
    t.r[5] = fma(bw1, w[2], a * w[0]);
    t.r[7] = fma(bw1, w[2], -a * w[0]);
    double wv[3], wwv[3];
    cross3(w, v, wv);
    cross3(w, wv, wwv);
#pragma unroll
    for (int i = 0; i < 3; i++)
        t.p[i] = fma(c2, wwv[i], fma(c1, wv[i], v[i]));
}

// SE3::log, include/liegroup.hpp:231-250.  Generic branch only while |ctheta| < 1; otherwise
// (theta = 0 and, as in the reference, also theta = pi) the result is (0, p).
//   theta = acos(c), sin(theta) = sqrt((1-c)(1+c)), 1/(2 tan(theta/2)) = (1+c)/(2 sin(theta))
//   w = vee(R - R^T) theta / (2 sin theta),  v = p - (w x p)/2 + coff1 (w x (w x p))
PPX_DEV void SE3_log(const Rigid &t, double (&xi)[6])
{
    const double ct = ((t.r[0] + t.r[4]) + t.r[8] - 1.0) * 0.5;
    if (fabs(ct) < 1.0)
    {
        const double theta = acos(ct);
        const double ist = 0.5 * rsqrt_nr((1.0 - ct) * (1.0 + ct)); // 1 / (2 sin theta); |ct| < 1 keeps the argument > 0
        const double ith = rcp_nr(theta);                          // theta >= acos(1 - 2^-53) ~ 1.5e-8
        const double k = theta * ist;
        const double coff1 = (ith - (1.0 + ct) * ist) * ith;
        const double w[3] = {(t.r[5] - t.r[7]) * k, (t.r[6] - t.r[2]) * k, (t.r[1] - t.r[3]) * k};
        double wp[3], wwp[3];
        cross3(w, t.p, wp);
        cross3(w, wp, wwp);
        xi[0] = w[0], xi[1] = w[1], xi[2] = w[2];
#pragma unroll
        for (int i = 0; i < 3; i++)
            xi[3 + i] = fma(coff1, wwp[i], fma(-0.5, wp[i], t.p[i]));
    }
    else
    {
        xi[0] = xi[1] = xi[2] = 0.0;
        xi[3] = t.p[0], xi[4] = t.p[1], xi[5] = t.p[2];
    }
}

// se3::adt, include/liegroup.hpp:271-280: [[hat(w), 0], [hat(v), hat(w)]] as a dense 6x6 record.
PPX_DEV void se3_adt_record(const double (&xi)[6], double (&ad)[36])
{
#pragma unroll
    for (int i = 0; i < 36; i++)
        ad[i] = 0.0;
    const double w0 = xi[0], w1 = xi[1], w2 = xi[2], v0 = xi[3], v1 = xi[4], v2 = xi[5];
    // hat(x) = [0 -x2 x1; x2 0 -x0; -x1 x0 0]
    ad[1 + 6 * 0] = w2, ad[2 + 6 * 0] = -w1, ad[0 + 6 * 1] = -w2, ad[2 + 6 * 1] = w0, ad[0 + 6 * 2] = w1, ad[1 + 6 * 2] = -w0;
    ad[4 + 6 * 0] = v2, ad[5 + 6 * 0] = -v1, ad[3 + 6 * 1] = -v2, ad[5 + 6 * 1] = v0, ad[3 + 6 * 2] = v1, ad[4 + 6 * 2] = -v0;
    ad[4 + 6 * 3] = w2, ad[5 + 6 * 3] = -w1, ad[3 + 6 * 4] = -w2, ad[5 + 6 * 4] = w0, ad[3 + 6 * 5] = w1, ad[4 + 6 * 5] = -w0;
}

// so3 ljac / ljacinv / rjac / rjacinv, include/liegroup.hpp:142-184 (which: 0..3).
//   ljac    = I + c1 X + c2 X^2         c1 = (1-cos x)/x^2, c2 = (x - sin x)/x^3
//   ljacinv = I - X/2 + c X^2           c  = 1/x^2 - (1+cos x)/(2 x sin x)
// with X = hat(w), X^2 = w w^T - x^2 I; below x^2 < EPS_DP the reference keeps I +- X/2.
PPX_DEV void so3_jac(int which, const double (&w)[3], double (&J)[9])
{
    const double x2 = fma(w[2], w[2], fma(w[1], w[1], w[0] * w[0]));
    const bool inv = which & 1;
    double lin, quad;
    if (x2 < EPS_DP)
    {
        lin = inv ? -0.5 : 0.5;
        quad = 0.0;
    }
    else
    {
        const double x = sqrt(x2);
        double s, c;
        sincos(x, &s, &c);
        if (!inv)
        {
            lin = (1.0 - c) / x2;
            quad = (x - s) / (x2 * x);
        }
        else
        {
            lin = -0.5;
            quad = 1.0 / x2 - (1.0 + c) / (2.0 * x * s);
        }
    }
    const double d = fma(-quad, x2, 1.0);
    double m[9];
    m[0] = fma(quad * w[0], w[0], d);
    m[4] = fma(quad * w[1], w[1], d);
    m[8] = fma(quad * w[2], w[2], d);
    m[1] = fma(quad * w[0], w[1], lin * w[2]);
    m[3] = fma(quad * w[0], w[1], -lin * w[2]);
    m[2] = fma(quad * w[0], w[2], -lin * w[1]);
    m[6] = fma(quad * w[0], w[2], lin * w[1]);
    m[5] = fma(quad * w[1], w[2], lin * w[0]);
    m[7] = fma(quad * w[1], w[2], -lin * w[0]);
    if (which >= 2) // rjac = ljac^T, rjacinv = ljacinv^T
    {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
                J[i + 3 * j] = m[j + 3 * i];
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 9; i++)
            J[i] = m[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Product-of-exponentials chain (demo/robotics.hpp).  The screws are the same for every instance,
// so everything that depends on them alone is prepared once per call on the host and handed to the
// kernel as a __grid_constant__ parameter: the compiler reads it from the constant bank as an FMA
// operand, it costs no registers and no per-thread arithmetic.
struct JointConst
{
    double w[3];    // omega
    double v[3];    // v
    double wv[3];   // omega x v
    double wwv[3];  // omega x (omega x v)
    double wn;      // |omega|
    double iw;      // 1/|omega|       (0 when omega = 0)
    double iw2;     // 1/|omega|^2     (0 when omega = 0)
    double iw3;     // 1/|omega|^3     (0 when omega = 0)
    double sgn;     // kind != 0: omega = sgn * e_(kind-1), sgn = +-1
    int kind;       // 0: general screw; 1, 2, 3: revolute about the x, y, z axis of the space frame (unit omega)
};

template <int NJ>
struct ChainConst
{
    JointConst joint[NJ];
    double home_r[9];
    double home_p[3];
};

// Host side: everything about the chain that does not depend on the instance (done once per call).
template <int NJ>
inline void fold_chain(const double *screws, const double *home, ChainConst<NJ> &ch)
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
        j.wn = sqrt(j.w[0] * j.w[0] + j.w[1] * j.w[1] + j.w[2] * j.w[2]);
        j.iw = j.wn > 0.0 ? 1.0 / j.wn : 0.0;
        j.iw2 = j.iw * j.iw;
        j.iw3 = j.iw2 * j.iw;
        // a unit rotation axis along a coordinate axis (the joints of every industrial arm in its home pose, all six of
        // the UR5's): the kernels then use the structure of that rotation instead of a dense 3x3 (poe_joint)
        j.kind = 0;
        j.sgn = 1.0;
        for (int a = 0; a < 3; a++)
            if (fabs(j.w[a]) == 1.0 && j.w[(a + 1) % 3] == 0.0 && j.w[(a + 2) % 3] == 0.0)
            {
                j.kind = a + 1;
                j.sgn = j.w[a];
            }
    }
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            ch.home_r[r + 3 * c] = home[r + 4 * c];
    ch.home_p[0] = home[12], ch.home_p[1] = home[13], ch.home_p[2] = home[14];
}

// exp(S q) for a fixed screw S = (w; v) and a per-instance angle q (se3::exp of `JList[i].screw * q`,
// demo/robotics.hpp:88-89).  With x = |w| q:
//   R = I + (sin x/|w|) W + ((1-cos x)/|w|^2) W^2,  p = q v + ((1-cos x)/|w|^2) (w x v) + ((x - sin x)/|w|^3) (w x (w x v))
// which is se3::exp's series with phi = |x|; no division and no square root per instance.
PPX_DEV void screw_exp(const JointConst &j, double q, Rigid &t)
{
    const double x = j.wn * q;
    double s, c;
    sincos(x, &s, &c);
    const bool tiny = fabs(x) < EPS_SP; // liegroup.hpp:258: R = I, p = v q
    const double a = tiny ? 0.0 : s * j.iw;
    const double b = tiny ? 0.0 : (1.0 - c) * j.iw2;
    const double e = tiny ? 0.0 : (x - s) * j.iw3;
    // W^2 = w w^T - |w|^2 I,  1 - b |w|^2 = cos(x)
    const double d = tiny ? 1.0 : c;
    const double bw0 = b * j.w[0], bw1 = b * j.w[1], bw2 = b * j.w[2];
    t.r[0] = fma(bw0, j.w[0], d);
    t.r[4] = fma(bw1, j.w[1], d);
    t.r[8] = fma(bw2, j.w[2], d);
    t.r[1] = fma(bw0, j.w[1], a * j.w[2]);
    t.r[3] = fma(bw0, j.w[1], -a * j.w[2]);
    t.r[2] = fma(bw0, j.w[2], -a * j.w[1]);
    t.r[6] = fma(bw0, j.w[2], a * j.w[1]);
    t.r[5] = fma(bw1, j.w[2], a * j.w[0]);
    t.r[7] = fma(bw1, j.w[2], -a * j.w[0]);
#pragma unroll
    for (int i = 0; i < 3; i++)
        t.p[i] = fma(e, j.wwv[i], fma(b, j.wv[i], q * j.v[i]));
}

// One joint of the product-of-exponentials chain: emits column i of the space Jacobian, Ad(P) S_i, from the prefix
// product P = e^{S_0 q_0} ... e^{S_{i-1} q_{i-1}} and then advances P <- P e^{S_i q_i}.
//
// KIND = 0: any screw (screw_exp + dense rigid product, ~98 fp64 instructions per joint).
// KIND = A + 1: omega = sgn e_A, a unit axis of the space frame.  exp(S q) then rotates by sgn*q ABOUT e_A, i.e. it
// leaves column A of P.R alone and mixes the other two with (cos, sin): 12 instructions instead of the dense 27, its
// rotation matrix is never formed, and R omega is sgn * column A -- ~54 instructions per joint.  Same formulas as the
// general path with the zeros and ones of omega taken out, so the results agree to the last bits (association of the
// sums differs) and the |x| < EPS_SP identity branch of se3::exp (liegroup.hpp:258-261) is the same select.
template <int KIND, class Emit>
PPX_DEV void poe_joint(const JointConst &j, double q, Rigid &P, int i, bool first, bool want_col, bool advance, Emit &&emit)
{
    if (KIND == 0)
    {
        if (want_col)
        {
            double col[6];
            if (first)
            {
#pragma unroll
                for (int k = 0; k < 3; k++)
                    col[k] = j.w[k], col[3 + k] = j.v[k];
            }
            else
                rigid_adjoint_apply(P, j.w, j.v, col);
            emit(i, col);
        }
        if (advance)
        {
            if (first)
                screw_exp(j, q, P);
            else
            {
                Rigid E, Q;
                screw_exp(j, q, E);
                rigid_mul(P, E, Q);
                P = Q;
            }
        }
        return;
    }
    constexpr int A = KIND > 0 ? KIND - 1 : 0, B = (A + 1) % 3, C = (A + 2) % 3;
    if (want_col)
    {
        double col[6];
        if (first)
        {
#pragma unroll
            for (int k = 0; k < 3; k++)
                col[k] = j.w[k], col[3 + k] = j.v[k];
        }
        else
        {
            // (R w ; p x (R w) + R v) with R w = sgn * column A of R
            const double rw[3] = {j.sgn * P.r[3 * A], j.sgn * P.r[3 * A + 1], j.sgn * P.r[3 * A + 2]};
            double rv[3], c[3];
            rot_apply(P.r, j.v, rv);
            cross3(P.p, rw, c);
            col[0] = rw[0], col[1] = rw[1], col[2] = rw[2];
            col[3] = c[0] + rv[0], col[4] = c[1] + rv[1], col[5] = c[2] + rv[2];
        }
        emit(i, col);
    }
    if (advance)
    {
        double s, c;
        sincos(q, &s, &c); // |omega| = 1: x = q
        const bool tiny = fabs(q) < EPS_SP;
        const double sn = tiny ? 0.0 : j.sgn * s; // rotation about +e_A by sgn*q
        const double cs = tiny ? 1.0 : c;
        const double b = tiny ? 0.0 : 1.0 - c;
        const double e = tiny ? 0.0 : q - s;
        double ep[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
            ep[k] = fma(e, j.wwv[k], fma(b, j.wv[k], q * j.v[k]));
        if (first)
        {
#pragma unroll
            for (int k = 0; k < 9; k++)
                P.r[k] = 0.0;
            P.r[A + 3 * A] = 1.0;
            P.r[B + 3 * B] = cs, P.r[C + 3 * B] = sn;   // column B = c e_B + s e_C
            P.r[B + 3 * C] = -sn, P.r[C + 3 * C] = cs;  // column C = -s e_B + c e_C
            P.p[0] = ep[0], P.p[1] = ep[1], P.p[2] = ep[2];
        }
        else
        {
            double np[3];
#pragma unroll
            for (int k = 0; k < 3; k++)
                np[k] = fma(P.r[k + 6], ep[2], fma(P.r[k + 3], ep[1], fma(P.r[k], ep[0], P.p[k])));
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const double xb = P.r[k + 3 * B], xc = P.r[k + 3 * C];
                P.r[k + 3 * B] = fma(sn, xc, cs * xb);
                P.r[k + 3 * C] = fma(cs, xc, -sn * xb);
                P.p[k] = np[k];
            }
        }
    }
}

// Where the kinds of a chain's joints come from.  RuntimeKinds reads them from the per-call constants: one uniform
// 4-way branch per joint, any chain.  FixedKinds<k0, k1, ...> makes them compile-time constants: the branch folds away
// and only the variant each joint needs is compiled -- used for the joint-axis patterns of the common 6R arm
// families (kinematics.cu: KindsUR, KindsWrist), where the branches and the 4x larger code cost the latency-bound IK
// kernels 8-17 % (profiles/r2r_ab.log); the memory-bound FK kernel is indifferent and always uses RuntimeKinds.
struct RuntimeKinds
{
    template <int NJ>
    PPX_DEV static int get(const ChainConst<NJ> &ch, int i)
    {
        return ch.joint[i].kind;
    }
};
template <int... K>
struct FixedKinds
{
    static constexpr int N = sizeof...(K);
    template <int NJ>
    PPX_DEV static int get(const ChainConst<NJ> &, int i)
    {
        constexpr int table[sizeof...(K) + 1] = {K..., 0};
        return table[i < N ? i : N];
    }
    // host side: does this chain have exactly these kinds?
    template <int NJ>
    static bool matches(const ChainConst<NJ> &ch)
    {
        constexpr int table[sizeof...(K) + 1] = {K..., 0};
        if (NJ != N)
            return false;
        for (int i = 0; i < NJ; i++)
            if (ch.joint[i].kind != table[i])
                return false;
        return true;
    }
};

// forwardSpace + jacobiSpace in one pass (demo/robotics.hpp:83-106).  The prefix products
// P_i = e^{S_0 q_0} ... e^{S_{i-1} q_{i-1}} are shared: column i of Js is Ad(P_i) S_i and the pose is
// P_N M.  (The reference associates forwardSpace right-to-left; sharing the prefixes changes the
// association, a 1e-16 effect.)  Each Jacobian column is handed to `emit(i, col)` as soon as it exists,
// so a caller that streams columns out never holds the 6 x NJ matrix in registers.  With RuntimeKinds the kind of every
// joint is a per-call constant (constant bank), so the switch below is a uniform branch: no divergence, and only the
// variant a chain uses is ever fetched; with FixedKinds it is resolved at compile time.
template <int NJ, bool WANT_T, bool WANT_J, class Kinds = RuntimeKinds, class Emit>
PPX_DEV void poe_fk_jacobian(const ChainConst<NJ> &ch, const double (&q)[NJ], Rigid &T, Emit &&emit)
{
    Rigid P;
#pragma unroll
    for (int i = 0; i < NJ; i++)
    {
        const bool first = i == 0, advance = WANT_T || i < NJ - 1;
        switch (Kinds::get(ch, i))
        {
        case 1: poe_joint<1>(ch.joint[i], q[i], P, i, first, WANT_J, advance, emit); break;
        case 2: poe_joint<2>(ch.joint[i], q[i], P, i, first, WANT_J, advance, emit); break;
        case 3: poe_joint<3>(ch.joint[i], q[i], P, i, first, WANT_J, advance, emit); break;
        default: poe_joint<0>(ch.joint[i], q[i], P, i, first, WANT_J, advance, emit); break;
        }
    }
    if (WANT_T)
    {
        Rigid M;
#pragma unroll
        for (int i = 0; i < 9; i++)
            M.r[i] = ch.home_r[i];
#pragma unroll
        for (int i = 0; i < 3; i++)
            M.p[i] = ch.home_p[i];
        rigid_mul(P, M, T);
    }
}

// ------------------------------------------------------------------------------------------------
// linsolve<Factorization::QR>, include/linalg.hpp:535-603, 1189-1200: Householder QR in the
// reference's (Numerical-Recipes qrdcmp) form, least-squares solve, SINGULAR flag when a pivot
// column has squared norm below EPS_DP; on SINGULAR x is the untouched head of b.
template <int M, int N>
PPX_DEV int linsolve_qr(double (&A)[M * N], double (&b)[M], double (&x)[N])
{
    static_assert(M >= N, "QR needs m >= n");
    double b0[N];
#pragma unroll
    for (int i = 0; i < N; i++)
        b0[i] = b[i];
    bool singular = false;
    double ic[N], id[N];
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        double sum = 0.0;
#pragma unroll
        for (int i = k; i < M; i++)
            sum = fma(A[i + k * M], A[i + k * M], sum);
        singular |= fabs(sum) < EPS_DP;
        const double root = sqrt(sum);
        const double sigma = A[k + k * M] >= 0.0 ? root : -root;
        A[k + k * M] += sigma;
        const double c = sigma * A[k + k * M];
        // 1/c and 1/d from one reciprocal of c*d (d = -sigma); tau is dropped when |c| <= EPS_DP (:564)
        const double rcd = 1.0 / (c * -sigma);
        ic[k] = fabs(c) > EPS_DP ? -sigma * rcd : 0.0;
        id[k] = c * rcd;
#pragma unroll
        for (int j = k + 1; j < N; j++)
        {
            double tau = 0.0;
#pragma unroll
            for (int i = k; i < M; i++)
                tau = fma(A[i + k * M], A[i + j * M], tau);
            tau *= ic[k];
#pragma unroll
            for (int i = k; i < M; i++)
                A[i + j * M] = fma(-tau, A[i + k * M], A[i + j * M]);
        }
        // apply the same reflector to b (QR::solve, :575-587)
        double s = 0.0;
#pragma unroll
        for (int i = k; i < M; i++)
            s = fma(A[i + k * M], b[i], s);
        s *= ic[k];
#pragma unroll
        for (int i = k; i < M; i++)
            b[i] = fma(-s, A[i + k * M], b[i]);
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--)
    {
        double s = b[i];
#pragma unroll
        for (int j = i + 1; j < N; j++)
            s = fma(-A[i + j * M], b[j], s);
        b[i] = s * id[i];
    }
#pragma unroll
    for (int i = 0; i < N; i++)
        x[i] = singular ? b0[i] : b[i];
    return singular ? 3 : 0;
}

// LU<n> with partial pivoting, include/linalg.hpp:448-533, in registers.  The pivot row is data dependent, so the
// row exchange is a predicated swap against every candidate row (compile-time indices only).  Pivot choice follows
// the reference: first row of maximal |A(row,k)|, SINGULAR as soon as that maximum is below EPS_DP (:473-477).
// The permutation is applied to the right-hand sides as it happens, which is what LU::solve's b[indx[row]] does.
// NRHS right-hand sides are carried along in B (column-major N x NRHS).  Returns the reference's `even` flag in
// `even`; on SINGULAR the factorisation stops being meaningful and the caller substitutes the reference's result.
// ipiv[k] receives 1 / u_kk.
template <int N, int NRHS>
PPX_DEV bool lu_factor_solve(double (&A)[N * N], double (&B)[N * (NRHS > 0 ? NRHS : 1)], bool &even, double (&ipiv)[N])
{
    bool singular = false;
    even = true;
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        double valmax = fabs(A[k + k * N]);
        int ip = k;
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const double t = fabs(A[r + k * N]);
            const bool better = valmax < t;
            valmax = better ? t : valmax;
            ip = better ? r : ip;
        }
        singular |= valmax < EPS_DP;
        even = (ip != k) ? !even : even;
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const bool sw = (ip == r);
#pragma unroll
            for (int c = 0; c < N; c++)
            {
                const double x = A[k + c * N], y = A[r + c * N];
                A[k + c * N] = sw ? y : x;
                A[r + c * N] = sw ? x : y;
            }
#pragma unroll
            for (int c = 0; c < NRHS; c++)
            {
                const double x = B[k + c * N], y = B[r + c * N];
                B[k + c * N] = sw ? y : x;
                B[r + c * N] = sw ? x : y;
            }
        }
        // one branch-free reciprocal per pivot serves the elimination and, later, the back substitution of every
        // right-hand side (the reference divides each time; the difference is one rounding)
        ipiv[k] = rcp_nr(A[k + k * N]);
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const double wgt = A[r + k * N] * ipiv[k];
            A[r + k * N] = wgt;
#pragma unroll
            for (int c = k + 1; c < N; c++)
                A[r + c * N] = fma(-wgt, A[k + c * N], A[r + c * N]);
#pragma unroll
            for (int c = 0; c < NRHS; c++) // forward substitution fused into the elimination
                B[r + c * N] = fma(-wgt, B[k + c * N], B[r + c * N]);
        }
    }
    // back substitution, :516-526
#pragma unroll
    for (int c = 0; c < NRHS; c++)
    {
#pragma unroll
        for (int r = N - 1; r >= 0; r--)
        {
            double sum = B[r + c * N];
#pragma unroll
            for (int j = r + 1; j < N; j++)
                sum = fma(-A[r + j * N], B[j + c * N], sum);
            B[r + c * N] = sum * ipiv[r];
        }
    }
    return singular;
}
template <int N, int NRHS>
PPX_DEV bool lu_factor_solve(double (&A)[N * N], double (&B)[N * (NRHS > 0 ? NRHS : 1)], bool &even)
{
    double ipiv[N];
    return lu_factor_solve<N, NRHS>(A, B, even, ipiv);
}

// Rigorous upper bound of cond_inf(A) from the partial-pivoting factors P A = L U held in `lu` (multipliers below the
// diagonal, |l_ij| <= 1; U on and above it; ipiv[k] = 1/u_kk):
//   |A|_inf <= |L|_inf |U|_inf <= N |U|_inf,     |A^-1|_inf <= |U^-1|_inf |L^-1|_inf <= |M(U)^-1 e|_inf 2^(N-1),
// where M(U) is U's comparison matrix (|u_ii| on the diagonal, -|u_ij| above it): M(U)^-1 >= |U^-1| entrywise, so one
// back substitution with absolute values bounds the norm of the inverse (Higham, "A survey of condition number
// estimation for triangular matrices", SIAM Rev. 29, 1987).  On 6x6 matrices the bound exceeds the true condition
// number by 10-300x (uniform random entries, UR5 Jacobians); zero, inf and NaN pivots give inf or NaN.
template <int N>
PPX_DEV double lu_cond_bound(const double (&lu)[N * N], const double (&ipiv)[N])
{
    double z[N], zmax = 0.0, unorm = 0.0;
#pragma unroll
    for (int i = N - 1; i >= 0; i--)
    {
        double acc = 1.0, row = fabs(lu[i + i * N]);
#pragma unroll
        for (int j = i + 1; j < N; j++)
        {
            acc = fma(fabs(lu[i + j * N]), z[j], acc);
            row += fabs(lu[i + j * N]);
        }
        z[i] = acc * fabs(ipiv[i]);
        zmax = z[i] > zmax ? z[i] : zmax; // a NaN z[i] leaves zmax alone but poisons every later z: caught via z[0]
        unorm = row > unorm ? row : unorm;
    }
    const double b = ((double)N * (double)(1 << (N - 1))) * unorm * zmax;
    return z[0] == z[0] ? b : z[0]; // NaN in, NaN out
}
// the LU shortcut of the square pseudo-inverse solve is taken below this bound: SVD::solve's truncation
// (0.5 sqrt(2n+1) eps w_max, linalg.hpp:889-901) acts only beyond cond_2 ~ 2.5e15 >> N * 1e10
constexpr double SQUARE_LU_COND_MAX = 1e10;

// MatrixS::I() (include/linalg.hpp:933-948: LU with partial pivoting, then one solve per unit vector; {} when SINGULAR)
// as an in-place Gauss-Jordan elimination with the same pivoting: the candidates for pivot k are the entries the LU
// sees (eliminating above the diagonal does not touch rows >= k), so the pivot sequence and the SINGULAR decision are
// the reference's, and the inverse overwrites A -- 36 live doubles for n = 6 instead of 72 for "factor + 6 right-hand
// sides", which is what lets this HBM-bound kernel run 4 CTAs per SM instead of 2.  Row exchanges are predicated swaps
// (compile-time indices only); they come back as column exchanges of the result, last first: A^-1 = (P A)^-1 P.
// pmin / pmax: smallest and largest pivot magnitude met (the conditioning guard of the square pinv, solvers.cu)
template <int N>
PPX_DEV bool gauss_jordan_inverse(double (&A)[N * N], double &pmin, double &pmax)
{
    bool singular = false;
    int piv[N];
    pmin = DBL_MAX;
    pmax = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        double valmax = fabs(A[k + k * N]);
        int ip = k;
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const double t = fabs(A[r + k * N]);
            const bool better = valmax < t;
            valmax = better ? t : valmax;
            ip = better ? r : ip;
        }
        singular |= valmax < EPS_DP;
        pmin = valmax < pmin ? valmax : pmin; // a NaN pivot leaves both untouched; the caller's test then fails on pmax
        pmax = valmax > pmax ? valmax : pmax;
        piv[k] = ip;
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const bool sw = (ip == r);
#pragma unroll
            for (int c = 0; c < N; c++)
            {
                const double x = A[k + c * N], y = A[r + c * N];
                A[k + c * N] = sw ? y : x;
                A[r + c * N] = sw ? x : y;
            }
        }
        const double d = rcp_nr(A[k + k * N]);
        A[k + k * N] = 1.0;
#pragma unroll
        for (int c = 0; c < N; c++)
            A[k + c * N] *= d;
#pragma unroll
        for (int r = 0; r < N; r++)
        {
            if (r == k)
                continue;
            const double f = A[r + k * N];
            A[r + k * N] = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++)
                A[r + c * N] = fma(-f, A[k + c * N], A[r + c * N]);
        }
    }
#pragma unroll
    for (int k = N - 2; k >= 0; k--)
    {
#pragma unroll
        for (int r = k + 1; r < N; r++)
        {
            const bool sw = (piv[k] == r);
#pragma unroll
            for (int i = 0; i < N; i++)
            {
                const double x = A[i + k * N], y = A[i + r * N];
                A[i + k * N] = sw ? y : x;
                A[i + r * N] = sw ? x : y;
            }
        }
    }
    return singular;
}
template <int N>
PPX_DEV bool gauss_jordan_inverse(double (&A)[N * N])
{
    double pmin, pmax;
    return gauss_jordan_inverse<N>(A, pmin, pmax);
}

// ------------------------------------------------------------------------------------------------
// Exact power-of-two pre-scaling for the Jacobi solvers, done entirely on the integer pipe (which is idle in these
// fp64-bound kernels).  The rotation formulas square differences of squared norms, i.e. they need |a|^4 to be
// representable; the reference's Householder steps divide by a running scale first (linalg.hpp:628-637, 982-996)
// and take any magnitude.  Subtracting `shift` from every exponent field brings the largest entry into [1, 2)
// exactly; entries more than 2^1022 below it flush to zero.  Results are scaled back the same way.
template <int L>
PPX_DEV int pow2_shift(const double (&a)[L])
{
    int m = 0;
#pragma unroll
    for (int i = 0; i < L; i++)
        m = max(m, __double2hiint(a[i]) & 0x7ff00000);
    // zero / denormal-only matrices and inf / NaN are left alone
    const int shift = (m == 0 || m == 0x7ff00000) ? 0 : (m >> 20) - 1023;
    // only magnitudes whose fourth power leaves the fp64 range need the treatment; the decision is taken per warp
    // so that the rescaling code is skipped (warp-uniformly) for ordinary data
    const bool need = shift > 200 || shift < -200;
    return __any_sync(0xffffffffu, need) ? shift : 0;
}
PPX_DEV double pow2_mul(double x, int shift) // x * 2^-shift for |shift| <= 1022, exponent arithmetic only
{
    const int hi = __double2hiint(x);
    const int f = (hi >> 20) & 0x7ff;
    const int nf = f - shift;
    const bool keep = f == 0 || f == 0x7ff; // zero, denormal, inf, NaN
    const int nhi = nf <= 0 ? (hi & 0x80000000) : (nf >= 0x7ff ? ((hi & 0x80000000) | 0x7ff00000) : hi - (shift << 20));
    const int nlo = (nf <= 0 || nf >= 0x7ff) ? 0 : __double2loint(x);
    return keep ? x : __hiloint2double(nhi, nlo);
}

template <int L>
PPX_DEV void pow2_rescale(double (&a)[L], int shift) // one warp-uniform branch around the whole array
{
    if (shift != 0)
    {
#pragma unroll
        for (int i = 0; i < L; i++)
            a[i] = pow2_mul(a[i], shift);
    }
}

// Jacobi rotation that annihilates the off-diagonal g of [[a, g], [g, b]]: with d = b - a, h = sqrt(d^2 + 4 g^2),
// u = |d| + h the tangent is t = sign(d) 2g / u, and  c = u / sqrt(u^2 + 4 g^2),  s = sign(d) 2g / sqrt(u^2 + 4 g^2)
// -- two reciprocal square roots, no division.
PPX_DEV void jacobi_cs(double a, double b, double g, double &c, double &s)
{
    const double d = b - a;
    const double g2 = g + g;
    const double gg = g2 * g2;
    const double r = fma(d, d, gg);
    const double h = r * rsqrt_nr(r);
    const double u = fabs(d) + h;
    const double w = rsqrt_nr(fma(u, u, gg));
    c = u * w;
    s = (d >= 0.0 ? g2 : -g2) * w;
}
// ... and the tangent as well (the two-sided update of the diagonal needs it); T = double, or float for the
// preconditioning sweeps
template <class T>
PPX_DEV void jacobi_cst(T a, T b, T g, T &c, T &s, T &t)
{
    const T d = b - a;
    const T g2 = g + g;
    const T gg = g2 * g2;
    const T r = t_fma(d, d, gg);
    const T h = r * rsqrt_nr(r);
    const T u = t_abs(d) + h;
    const T w = rsqrt_nr(t_fma(u, u, gg));
    const T sg = d >= T(0) ? g2 : -g2;
    c = u * w;
    s = sg * w;
    t = sg * rcp_nr(u);
}

// Round-robin ("tournament") ordering of the N(N-1)/2 index pairs: ROUNDS rounds of SLOTS mutually
// disjoint pairs (circle method; odd N plays against a dummy index N that is skipped).  Disjoint pairs touch
// different columns, so the long dependent chains that produce each rotation (dot products, square root,
// division, reciprocal square root) are independent within a round and the compiler interleaves them --
// that instruction-level parallelism is what keeps the fp64 pipe busy at 2 warps per scheduler.
template <int N>
struct Tournament
{
    static constexpr int E = N + (N & 1); // even number of players
    static constexpr int ROUNDS = E - 1;
    static constexpr int SLOTS = E / 2;
    __host__ __device__ static constexpr int first(int r, int k) { return k == 0 ? E - 1 : (r + k) % (E - 1); }
    __host__ __device__ static constexpr int second(int r, int k) { return k == 0 ? r : (r - k + (E - 1)) % (E - 1); }
    __host__ __device__ static constexpr int lo(int r, int k) { return first(r, k) < second(r, k) ? first(r, k) : second(r, k); }
    __host__ __device__ static constexpr int hi(int r, int k) { return first(r, k) < second(r, k) ? second(r, k) : first(r, k); }
    __host__ __device__ static constexpr bool valid(int r, int k) { return hi(r, k) < N; }
};

// cos^2 of the angle between two columns below which a rotation counts as "fine" (see jacobi_svd_core's look-ahead)
constexpr double SVD_FINE_COS2 = 1e-10;

// One rotation of the one-sided Jacobi SVD on the columns at positions p and p+1, derived in two phases so that
// the rotations of a round (disjoint positions) can be derived together: their chains are independent and overlap.
template <int M, class T>
PPX_DEV void svd_pair_derive(const T (&ap)[M], const T (&aq)[M], T &np_, T &nq_, T tol2, T fine2, T &c, T &s, bool &act,
                             bool &coarse)
{
    T ga = T(0);
#pragma unroll
    for (int i = 0; i < M; i++)
        ga = t_fma(ap[i], aq[i], ga);
    const T g2 = ga * ga, nn = np_ * nq_;
    act = g2 > tol2 * nn;
    coarse |= g2 > fine2 * nn;
    T t;
    jacobi_cst<T>(np_, nq_, ga, c, s, t);
    c = act ? c : T(1);
    s = act ? s : T(0);
    t = act ? t : T(0);
    // the rotated columns SWAP positions (odd-even transposition ordering): so do their norms
    const T n0 = t_fma(-t, ga, np_), n1 = t_fma(t, ga, nq_);
    np_ = n1;
    nq_ = n0;
}

// x' = c x - s y,  y' = s x + c y, stored swapped: position p gets y', position p+1 gets x'
template <int L, class T>
PPX_DEV void rotate_swap(T (&xp)[L], T (&xq)[L], T c, T s)
{
#pragma unroll
    for (int i = 0; i < L; i++)
    {
        const T x = xp[i], y = xq[i];
        xp[i] = t_fma(s, x, c * y);
        xq[i] = t_fma(c, x, -s * y);
    }
}

template <int L>
PPX_DEV void swap_cols(double (&xp)[L], double (&xq)[L])
{
#pragma unroll
    for (int i = 0; i < L; i++)
    {
        const double x = xp[i];
        xp[i] = xq[i];
        xq[i] = x;
    }
}

// One round of the odd-even ordering: positions (START, START+1), (START+2, START+3), ...
template <int M, int N, int START, bool WANT_V, bool WANT_C, class T>
PPX_DEV void svd_round(T (&a)[N][M], T (&v)[N][N], T (&cv)[N], T (&nrm)[N], T tol2, T fine2, bool &rotated, bool &coarse)
{
    constexpr int P = (N - START) / 2;
    T c[P > 0 ? P : 1], s[P > 0 ? P : 1];
    bool act[P > 0 ? P : 1];
#pragma unroll
    for (int k = 0; k < P; k++)
    {
        svd_pair_derive<M, T>(a[START + 2 * k], a[START + 2 * k + 1], nrm[START + 2 * k], nrm[START + 2 * k + 1], tol2,
                              fine2, c[k], s[k], act[k], coarse);
        rotated |= act[k];
    }
#pragma unroll
    for (int k = 0; k < P; k++)
    {
        // always applied, also when no lane needs it (c = 1, s = 0 is then the plain swap): a warp-uniform skip
        // would save the multiply-adds of the last sweep but costs register moves at every merge point of the
        // rolled loop -- measured slower
        rotate_swap<M, T>(a[START + 2 * k], a[START + 2 * k + 1], c[k], s[k]);
        if (WANT_V)
            rotate_swap<N, T>(v[START + 2 * k], v[START + 2 * k + 1], c[k], s[k]);
        if (WANT_C) // a vector with one entry per column, carried through the same rotations
        {
            const T x = cv[START + 2 * k], y = cv[START + 2 * k + 1];
            cv[START + 2 * k] = t_fma(s[k], x, c[k] * y);
            cv[START + 2 * k + 1] = t_fma(c[k], x, -s[k] * y);
        }
    }
}

// ---- packed single precision (sm_100: FFMA2 / FMUL2 do two fp32 operations per issue slot) ----------------------
#ifdef __CUDA_ARCH__
typedef float2 F2;
PPX_DEV F2 f2_fma(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
PPX_DEV F2 f2_mul(F2 a, F2 b) { return __fmul2_rn(a, b); }
#else
struct F2
{
    float x, y;
};
PPX_DEV F2 f2_fma(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
PPX_DEV F2 f2_mul(F2 a, F2 b) { return F2{a.x * b.x, a.y * b.y}; }
#endif
PPX_DEV F2 f2_set(float x, float y)
{
    F2 r;
    r.x = x, r.y = y;
    return r;
}

// Single-precision preconditioner of the Jacobi SVD.  The fp64 pipe is what bounds these kernels (64 DFMA per clock
// and SM against 128 FFMA, and a packed FFMA2 does two of those per issue slot), and most of a Jacobi SVD's arithmetic
// is spent getting NEAR the answer: the first 3-4 of its ~5 sweeps.  So those are run on a float copy of the matrix
// (same odd-even ordering, fixed even number of sweeps, no votes; rows held in pairs so that dot products and
// rotations are FFMA2 / FMUL2), only to produce the accumulated rotations V0 ~ the right singular vectors to ~1e-6.
// V0 is then made orthonormal to fp64 accuracy by modified Gram-Schmidt (it is nearly orthogonal, so MGS loses
// nothing), and the fp64 iteration starts from A V0 with V = V0: 2 sweeps instead of 5 (measured on the host build,
// tools/svd_sweeps.cpp).  Every number that leaves the kernel is computed by the fp64 iteration with its own
// convergence test -- the preconditioner can only change how many sweeps that takes, not what it converges to.
// (fp32 range: the copy is scaled by an exact power of two so that the largest entry is in [1, 2); entries 2^-126
// below it flush to zero, which a preconditioner may ignore.)
template <int MH, int NH>
PPX_DEV void f32_rotate_pair(F2 (&ap)[MH], F2 (&aq)[MH], F2 (&vp)[NH], F2 (&vq)[NH], float &np_, float &nq_)
{
    F2 acc = f2_set(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < MH; i++)
        acc = f2_fma(ap[i], aq[i], acc);
    const float ga = acc.x + acc.y;
    // the rotation that annihilates ga (jacobi_cst), without the "is it worth rotating" test of the fp64 iteration:
    // the sweep count is fixed here, and a negligible ga gives c = 1, s = t = 0 by itself once the radicand cannot be
    // zero (1e-37 is below the square of anything the scaled copy can hold next to an entry in [1, 2))
    const float d = nq_ - np_;
    const float g2 = ga + ga;
    const float gg = g2 * g2;
    const float r = fmaf(d, d, gg) + 1e-37f;
    const float h = r * rsqrt_nr(r);
    const float u = fabsf(d) + h;
    const float w = rsqrt_nr(fmaf(u, u, gg));
    const float sg = d >= 0.0f ? g2 : -g2;
    const float c = u * w, s = sg * w, t = sg * rcp_nr(u);
    const float n0 = fmaf(-t, ga, np_), n1 = fmaf(t, ga, nq_);
    np_ = n1; // the rotated columns swap positions
    nq_ = n0;
    const F2 c2 = f2_set(c, c), s2 = f2_set(s, s), ns2 = f2_set(-s, -s);
#pragma unroll
    for (int i = 0; i < MH; i++)
    {
        const F2 x = ap[i], y = aq[i];
        ap[i] = f2_fma(s2, x, f2_mul(c2, y));
        aq[i] = f2_fma(c2, x, f2_mul(ns2, y));
    }
#pragma unroll
    for (int i = 0; i < NH; i++)
    {
        const F2 x = vp[i], y = vq[i];
        vp[i] = f2_fma(s2, x, f2_mul(c2, y));
        vq[i] = f2_fma(c2, x, f2_mul(ns2, y));
    }
}

template <int M, int N>
PPX_DEV void jacobi_precondition_f32(const double (&a64)[N][M], double (&v0)[N][N])
{
#ifndef PPX_SVD_F32_SWEEPS
#define PPX_SVD_F32_SWEEPS 4
#endif
    constexpr int SWEEPS = PPX_SVD_F32_SWEEPS; // even: the odd-even ordering reverses the column order every sweep
    static_assert(SWEEPS % 2 == 0, "an odd number of sweeps would leave the columns reversed");
    constexpr int MH = (M + 1) / 2, NH = (N + 1) / 2; // rows in pairs; an odd count is padded with a zero row
    int mexp = 0;
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < M; i++)
            mexp = max(mexp, __double2hiint(a64[j][i]) & 0x7ff00000);
    // 2^-(e) with e = exponent of the largest entry; left at 1 for zero / denormal-only / inf / NaN matrices
    const int e = (mexp >> 20) - 1023;
    const bool scalable = mexp != 0 && mexp != 0x7ff00000 && e > -1000 && e < 1000;
    const double sc = scalable ? __hiloint2double((1023 - e) << 20, 0) : 1.0;
    F2 a[N][MH], v[N][NH];
    float nrm[N];
#pragma unroll
    for (int j = 0; j < N; j++)
    {
#pragma unroll
        for (int i = 0; i < MH; i++)
            a[j][i] = f2_set((float)(a64[j][2 * i] * sc), 2 * i + 1 < M ? (float)(a64[j][2 * i + 1 < M ? 2 * i + 1 : 0] * sc) : 0.0f);
#pragma unroll
        for (int i = 0; i < NH; i++)
            v[j][i] = f2_set(2 * i == j ? 1.0f : 0.0f, 2 * i + 1 == j ? 1.0f : 0.0f);
    }
#pragma unroll 1
    for (int sweep = 0; sweep < SWEEPS; sweep++)
    {
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            F2 acc = f2_set(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < MH; i++)
                acc = f2_fma(a[j][i], a[j][i], acc);
            nrm[j] = acc.x + acc.y;
        }
#pragma unroll 1
        for (int it = 0; it < N / 2; it++)
        {
#pragma unroll
            for (int k = 0; k + 1 < N; k += 2)
                f32_rotate_pair<MH, NH>(a[k], a[k + 1], v[k], v[k + 1], nrm[k], nrm[k + 1]);
#pragma unroll
            for (int k = 1; k + 1 < N; k += 2)
                f32_rotate_pair<MH, NH>(a[k], a[k + 1], v[k], v[k + 1], nrm[k], nrm[k + 1]);
        }
        if (N & 1)
        {
#pragma unroll
            for (int k = 0; k + 1 < N; k += 2)
                f32_rotate_pair<MH, NH>(a[k], a[k + 1], v[k], v[k + 1], nrm[k], nrm[k + 1]);
        }
    }
    // V0 in fp64, orthonormalised by modified Gram-Schmidt (an even number of sweeps left the columns in order)
#pragma unroll
    for (int j = 0; j < N; j++)
    {
#pragma unroll
        for (int i = 0; i < N; i++)
            v0[j][i] = (double)((i & 1) ? v[j][i / 2].y : v[j][i / 2].x);
#pragma unroll
        for (int k = 0; k < j; k++)
        {
            double r = 0.0;
#pragma unroll
            for (int i = 0; i < N; i++)
                r = fma(v0[k][i], v0[j][i], r);
#pragma unroll
            for (int i = 0; i < N; i++)
                v0[j][i] = fma(-r, v0[k][i], v0[j][i]);
        }
        double n2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++)
            n2 = fma(v0[j][i], v0[j][i], n2);
        // a NaN / inf input poisons the float sweeps: fall back to the identity column rather than spreading NaN
        const bool good = n2 > 0.25 && n2 < 4.0;
        const double inv = rsqrt_nr(good ? n2 : 1.0);
#pragma unroll
        for (int i = 0; i < N; i++)
            v0[j][i] = good ? v0[j][i] * inv : ((i == j) ? 1.0 : 0.0);
    }
}

// One-sided (Hestenes) Jacobi SVD of an M x N matrix held column-major in registers:
// on return A holds U diag(w) (columns = w_j u_j), w2[j] = |column j|^2, V the accumulated rotations.
//
// Odd-even transposition ordering (Luk & Park): even rounds rotate the neighbouring positions (0,1), (2,3), ...,
// odd rounds (1,2), (3,4), ..., and every rotation leaves its two columns SWAPPED.  After N rounds every pair of
// columns has met exactly once and the column order is reversed.  The swap is free -- the rotation just writes its
// two results to the opposite registers -- so, unlike a round-robin schedule, every round runs the same code on
// the same registers with no data movement at all, and the sweep is a rolled loop over (even round, odd round):
// ~8 KB of SASS instead of ~25 KB for an unrolled sweep, which is what the instruction cache needs at 2 warps
// per scheduler (ncu on the unrolled form: more warp-cycles stalled on `no_instruction` than issuing).
//
// A pair is rotated while |a_p . a_q| > tol |a_p| |a_q| with tol = 4 sqrt(M) eps (LAPACK's dgesvj uses
// sqrt(M) eps; the singular values come out with relative accuracy ~ tol).  Sweeps stop when a whole sweep
// rotated nothing in any lane of the warp.  Squared column norms are recomputed at the start of every sweep and
// carried through its rotations (a rotation by tangent t moves t*gamma from one norm to the other).
#ifdef PPX_SVD_COUNT_SWEEPS
static long g_ppx_svd_sweeps = 0;
#endif
#ifndef PPX_SVD_PRECOND
#define PPX_SVD_PRECOND 1 // single-precision preconditioning sweeps (jacobi_precondition_f32); 0: fp64 from the start
#endif
// PRECOND: run the single-precision preconditioning sweeps first.  It pays where V is wanted anyway (SVD, pinv,
// rectangular linsolve<SVD>: +23-25 %) and for norm2 (+12 %); the row-orthogonalising square solve, which otherwise
// accumulates no V at all, loses with it (IK step with the SVD solve forced: 1.84 against 2.13 G inst/s).
template <int M, int N, bool WANT_V, bool WANT_C = false, bool PRECOND = (PPX_SVD_PRECOND != 0 && !WANT_C)>
PPX_DEV void jacobi_svd_core(double (&A)[M * N], double (&V)[N * N], double (&w2)[N], double *carry = nullptr)
{
    double a[N][M], v[N][N], nrm[N], cv[N];
#pragma unroll
    for (int j = 0; j < N; j++)
        cv[j] = WANT_C ? carry[j] : 0.0;
#pragma unroll
    for (int j = 0; j < N; j++)
    {
#pragma unroll
        for (int i = 0; i < M; i++)
            a[j][i] = A[i + j * M];
#pragma unroll
        for (int i = 0; i < N; i++)
            v[j][i] = (i == j) ? 1.0 : 0.0;
    }
    if (PRECOND && N >= 3)
    {
        // start the fp64 iteration from A V0, V = V0 (and the carried vector from V0^T carry)
        jacobi_precondition_f32<M, N>(a, v);
        double t[N][M];
#pragma unroll
        for (int j = 0; j < N; j++)
#pragma unroll
            for (int i = 0; i < M; i++)
            {
                double acc = a[0][i] * v[j][0];
#pragma unroll
                for (int k = 1; k < N; k++)
                    acc = fma(a[k][i], v[j][k], acc);
                t[j][i] = acc;
            }
        if (WANT_C)
        {
            double tc[N];
#pragma unroll
            for (int j = 0; j < N; j++)
            {
                double acc = cv[0] * v[j][0];
#pragma unroll
                for (int k = 1; k < N; k++)
                    acc = fma(cv[k], v[j][k], acc);
                tc[j] = acc;
            }
#pragma unroll
            for (int j = 0; j < N; j++)
                cv[j] = tc[j];
        }
#pragma unroll
        for (int j = 0; j < N; j++)
#pragma unroll
            for (int i = 0; i < M; i++)
                a[j][i] = t[j][i];
    }
    constexpr double TOL2 = 16.0 * M * EPS_DP * EPS_DP;
    constexpr int MAX_SWEEPS = 12;
    bool reversed = false;
    for (int sweep = 0; sweep < MAX_SWEEPS; sweep++)
    {
#pragma unroll
        for (int j = 0; j < N; j++)
        {
            double a2 = 0.0;
#pragma unroll
            for (int i = 0; i < M; i++)
                a2 = fma(a[j][i], a[j][i], a2);
            nrm[j] = a2;
        }
        bool rotated = false, coarse = false;
#pragma unroll 1
        for (int it = 0; it < N / 2; it++)
        {
            svd_round<M, N, 0, WANT_V, WANT_C, double>(a, v, cv, nrm, TOL2, SVD_FINE_COS2, rotated, coarse);
            svd_round<M, N, 1, WANT_V, WANT_C, double>(a, v, cv, nrm, TOL2, SVD_FINE_COS2, rotated, coarse);
        }
        if (N & 1)
            svd_round<M, N, 0, WANT_V, WANT_C, double>(a, v, cv, nrm, TOL2, SVD_FINE_COS2, rotated, coarse);
        reversed = !reversed;
#ifdef PPX_SVD_COUNT_SWEEPS // host-side instrumentation of tests/cpp (never defined for the device build)
        g_ppx_svd_sweeps++;
#endif
        if (!__any_sync(0xffffffffu, rotated))
            break;
#ifndef PPX_SVD_NO_LOOKAHEAD
        // Look-ahead.  Convergence is quadratic, so a sweep whose rotations were all fine (cos < 1e-5) is nearly
        // always the last one that rotates, and the sweep after it would only establish that: N(N-1)/2 pairs derived
        // and "rotated" by the identity.  The same test -- every pair within tolerance, norms fresh -- costs a quarter
        // of that when made directly.  It is exactly the decision the next sweep would take (no rotation happens in
        // it, so the order in which it meets the pairs does not matter), hence the results are bit-identical.
        if (!__any_sync(0xffffffffu, coarse))
        {
            double n2[N];
#pragma unroll
            for (int j = 0; j < N; j++)
            {
                double a2 = 0.0;
#pragma unroll
                for (int i = 0; i < M; i++)
                    a2 = fma(a[j][i], a[j][i], a2);
                n2[j] = a2;
            }
            bool clean = true;
#pragma unroll
            for (int p = 0; p < N; p++)
#pragma unroll
                for (int q = p + 1; q < N; q++)
                {
                    double ga = 0.0;
#pragma unroll
                    for (int i = 0; i < M; i++)
                        ga = fma(a[p][i], a[q][i], ga);
                    clean = clean && !(ga * ga > TOL2 * (n2[p] * n2[q]));
                }
            if (__all_sync(0xffffffffu, clean))
                break;
        }
#endif
    }
    // an odd number of sweeps leaves the columns in reverse order: undo it (the sweep count is warp-uniform)
#pragma unroll
    for (int j = 0; j < N; j++)
    {
        double s2 = 0.0;
#pragma unroll
        for (int i = 0; i < M; i++)
        {
            const double x = reversed ? a[N - 1 - j][i] : a[j][i];
            A[i + j * M] = x;
            s2 = fma(x, x, s2);
        }
        w2[j] = s2;
        if (WANT_V)
        {
#pragma unroll
            for (int i = 0; i < N; i++)
                V[i + j * N] = reversed ? v[N - 1 - j][i] : v[j][i];
        }
        if (WANT_C)
            carry[j] = reversed ? cv[N - 1 - j] : cv[j];
    }
}

// Truncated pseudo-inverse solve from the one-sided Jacobi factors (SVD::solve, linalg.hpp:886-913):
// x = sum_j v_j (a_j . b) / w_j^2 over w_j > tsh = 0.5 sqrt(m+n+1) w_max EPS_DP, where a_j = w_j u_j.
template <int M, int N>
PPX_DEV void jacobi_svd_solve(const double (&A)[M * N], const double (&V)[N * N], const double (&w2)[N],
                              const double (&b)[M], double (&x)[N])
{
    double wmax2 = w2[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        wmax2 = fmax(wmax2, w2[j]);
    const double f = 0.5 * sqrt((double)(M + N + 1)) * EPS_DP;
    const double tsh2 = f * f * wmax2;
#pragma unroll
    for (int i = 0; i < N; i++)
        x[i] = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++)
    {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < M; i++)
            s = fma(A[i + j * M], b[i], s);
        const double coef = w2[j] > tsh2 ? s / w2[j] : 0.0;
#pragma unroll
        for (int i = 0; i < N; i++)
            x[i] = fma(V[i + j * N], coef, x[i]);
    }
}

// Truncated pseudo-inverse solve of a SQUARE system without accumulating V (or U): orthogonalise the ROWS of A with the
// same Jacobi machinery (columns of A^T) and carry b through the rotations.  With G = U^T A (orthogonal rows g_j of
// norm w_j) and c = U^T b,  A^+ b = G^T diag(1/w^2) c = sum_j g_j c_j / w_j^2  over w_j > tsh -- the same truncation as
// SVD::solve (linalg.hpp:886-913), 64 instead of 88 fp64 instructions per rotation.
template <int N, bool PRECOND = false>
PPX_DEV void jacobi_rowsolve(const double (&A)[N * N], const double (&b)[N], double (&x)[N])
{
    double At[N * N], dummy[N * N], w2[N], c[N];
#pragma unroll
    for (int r = 0; r < N; r++)
    {
        c[r] = b[r];
#pragma unroll
        for (int j = 0; j < N; j++)
            At[j + r * N] = A[r + j * N];
    }
    jacobi_svd_core<N, N, false, true, PRECOND>(At, dummy, w2, c);
    double wmax2 = w2[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        wmax2 = fmax(wmax2, w2[j]);
    const double f = 0.5 * sqrt((double)(2 * N + 1)) * EPS_DP;
    const double tsh2 = f * f * wmax2;
#pragma unroll
    for (int i = 0; i < N; i++)
        x[i] = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++)
    {
        const double coef = w2[j] > tsh2 ? c[j] / w2[j] : 0.0;
#pragma unroll
        for (int i = 0; i < N; i++)
            x[i] = fma(At[i + j * N], coef, x[i]);
    }
}

// linsolve<Factorization::SVD> on a SQUARE system (SVD::solve, linalg.hpp:886-913) = A^+ b with singular values below
// 0.5 sqrt(2n+1) w_max EPS_DP dropped.  That truncation can only act when cond(A) > ~2e15; everywhere else A^+ b is
// A^-1 b, and a backward-stable solver returns it to the same accuracy (~cond eps) whatever the factorisation.  So the
// system is first solved by LU with partial pivoting (~1/30 of the arithmetic of the Jacobi solve) and a rigorous upper
// bound of cond_inf(A) is read off the factors (lu_cond_bound -- pivot ratios alone do not bound the condition number:
// I + 1e3 triu(ones) has unit pivots and cond 3.5e18).  Unless that bound is below 1e10 -- five orders of magnitude
// before truncation could matter; false for zero, inf and NaN pivots -- the row-orthogonalising Jacobi solve is run as
// well and its result taken for those lanes.  The decision to run it is per warp, so well-conditioned batches never
// execute it.  `allow_lu = false` forces the Jacobi path.
// (Tried and dropped: the Jacobi fallback as a __noinline__ function, so that the common path would not share its
// register allocation -- the registers live across the call get spilled around it and the common path pays: IK Newton
// step 8.7 against 10.0 G inst/s.  A separate second-pass kernel over the instances the LU pass marks was measured too:
// 7.9-8.3 G inst/s.  The fallback stays inline, behind a warp vote.)
template <int N, bool PRECOND_FALLBACK = false>
PPX_DEV void square_pinv_solve(const double (&A)[N * N], const double (&b)[N], double (&x)[N], bool allow_lu)
{
    double rx[N];
#pragma unroll
    for (int i = 0; i < N; i++)
        rx[i] = b[i];
    bool easy = false;
    if (allow_lu) // uniform over the launch
    {
        double lu[N * N];
#pragma unroll
        for (int i = 0; i < N * N; i++)
            lu[i] = A[i];
        bool even;
        double ipiv[N];
        lu_factor_solve<N, 1>(lu, rx, even, ipiv);
        easy = lu_cond_bound<N>(lu, ipiv) < SQUARE_LU_COND_MAX;
        if (__all_sync(0xffffffffu, easy))
        {
#pragma unroll
            for (int i = 0; i < N; i++)
                x[i] = rx[i];
            return;
        }
    }
    double xs[N];
    jacobi_rowsolve<N, PRECOND_FALLBACK>(A, b, xs);
#pragma unroll
    for (int i = 0; i < N; i++)
        x[i] = easy ? rx[i] : xs[i];
}

// Cyclic Jacobi eigensolver for a symmetric N x N matrix (EigenValue<n>, include/linalg.hpp:967-1167:
// same result as tred2 + implicit QL up to rounding and eigenvector sign).  S is read through its
// lower triangle like the reference's tred2 (z(i,k), k <= i).  d = eigenvalues, Z = eigenvectors in
// columns; `sorted` reproduces eigsrt's ascending order (:1137-1151).
// index of the (i, j) entry of a symmetric matrix kept as its upper triangle in a dense N*N array
__host__ __device__ constexpr int sym_ix(int i, int j, int n) { return i <= j ? i + j * n : j + i * n; }

template <int N, bool WANT_Z>
PPX_DEV void jacobi_eig_sym(const double (&S)[N * N], bool sorted, double (&d)[N], double (&Z)[N * N])
{
    double a[N * N]; // only the entries sym_ix() maps to are live
    double fro2 = 0.0;
    const int shift = pow2_shift(S);
    double Ss[N * N];
#pragma unroll
    for (int i = 0; i < N * N; i++)
        Ss[i] = S[i];
    pow2_rescale(Ss, shift);
#pragma unroll
    for (int c = 0; c < N; c++)
#pragma unroll
        for (int r = c; r < N; r++)
        {
            const double v = Ss[r + c * N]; // lower triangle, as tred2 reads it
            a[sym_ix(r, c, N)] = v;
            fro2 = fma(r == c ? v : 2.0 * v, v, fro2);
        }
    if (WANT_Z)
    {
#pragma unroll
        for (int i = 0; i < N * N; i++)
            Z[i] = (i % (N + 1) == 0) ? 1.0 : 0.0;
    }
    // an off-diagonal is negligible once |g| <= eps |S|_F / 2: the eigenvalues then carry an absolute error of
    // a few eps |S|_2, the same guarantee the reference's QL iteration gives (|e| < eps (|d_m| + |d_m+1|))
    const double thr2 = 0.25 * EPS_DP * EPS_DP * fro2;
    using TT = Tournament<N>;
    constexpr int MAX_SWEEPS = 12;
    for (int sweep = 0; sweep < MAX_SWEEPS; sweep++)
    {
        // Done when no off-diagonal of any lane is above the threshold.  Testing that directly costs N(N-1)/2 multiplies
        // and compares; finding it out by running a sweep in which no rotation acts costs that sweep's N(N-1)/2 rotation
        // derivations (two reciprocal square roots and a reciprocal each) -- it is the same decision, so the results are
        // bit-identical, and the last sweep of every matrix is saved
        bool big = false;
#pragma unroll
        for (int p = 0; p < N; p++)
#pragma unroll
            for (int q = p + 1; q < N; q++)
                big |= a[sym_ix(p, q, N)] * a[sym_ix(p, q, N)] > thr2;
        if (!__any_sync(0xffffffffu, big))
            break;
        bool rotated = false;
#pragma unroll
        for (int r = 0; r < TT::ROUNDS; r++)
        {
            // the pairs of a round are disjoint: rotating (p,q) leaves a_p'p', a_q'q', a_p'q' of the other pairs
            // untouched, so all rotations of the round can be derived first (independent chains) and applied after
            double c[TT::SLOTS], s[TT::SLOTS], t[TT::SLOTS];
            bool act[TT::SLOTS];
#pragma unroll
            for (int k = 0; k < TT::SLOTS; k++)
            {
                if (!TT::valid(r, k))
                    continue;
                const int p = TT::lo(r, k), q = TT::hi(r, k);
                const double g = a[sym_ix(p, q, N)];
                act[k] = g * g > thr2;
                jacobi_cst(a[sym_ix(p, p, N)], a[sym_ix(q, q, N)], g, c[k], s[k], t[k]);
                c[k] = act[k] ? c[k] : 1.0;
                s[k] = act[k] ? s[k] : 0.0;
                t[k] = act[k] ? t[k] : 0.0;
                rotated |= act[k];
            }
#pragma unroll
            for (int k = 0; k < TT::SLOTS; k++)
            {
                if (!TT::valid(r, k))
                    continue;
                const int p = TT::lo(r, k), q = TT::hi(r, k);
                if (__any_sync(0xffffffffu, act[k]))
                {
                    const double g = a[sym_ix(p, q, N)];
                    a[sym_ix(p, p, N)] = fma(-t[k], g, a[sym_ix(p, p, N)]);
                    a[sym_ix(q, q, N)] = fma(t[k], g, a[sym_ix(q, q, N)]);
                    a[sym_ix(p, q, N)] = act[k] ? 0.0 : g;
#pragma unroll
                    for (int i = 0; i < N; i++)
                    {
                        if (i != p && i != q)
                        {
                            const double x = a[sym_ix(i, p, N)], y = a[sym_ix(i, q, N)];
                            a[sym_ix(i, p, N)] = fma(c[k], x, -s[k] * y);
                            a[sym_ix(i, q, N)] = fma(s[k], x, c[k] * y);
                        }
                    }
                    if (WANT_Z)
                    {
#pragma unroll
                        for (int i = 0; i < N; i++)
                        {
                            const double x = Z[i + p * N], y = Z[i + q * N];
                            Z[i + p * N] = fma(c[k], x, -s[k] * y);
                            Z[i + q * N] = fma(s[k], x, c[k] * y);
                        }
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, rotated))
            break;
    }
#pragma unroll
    for (int i = 0; i < N; i++)
        d[i] = a[i + i * N];
    pow2_rescale(d, -shift);
    if (sorted)
    {
        // ascending selection sort with compile-time indices (eigsrt swaps whole columns)
#pragma unroll
        for (int i = 0; i < N - 1; i++)
#pragma unroll
            for (int j = i + 1; j < N; j++)
            {
                const bool sw = d[j] < d[i];
                const double di = d[i], dj = d[j];
                d[i] = sw ? dj : di;
                d[j] = sw ? di : dj;
                if (WANT_Z)
                {
#pragma unroll
                    for (int r = 0; r < N; r++)
                    {
                        const double x = Z[r + i * N], y = Z[r + j * N];
                        Z[r + i * N] = sw ? y : x;
                        Z[r + j * N] = sw ? x : y;
                    }
                }
            }
    }
}

} // namespace ppx_dev
