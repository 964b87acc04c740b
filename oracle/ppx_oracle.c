/*
 * ppx_oracle.c -- plain-C restatement of the reference's hot-path algorithms.
 *
 * TEST INFRASTRUCTURE ONLY (see ppx_oracle.h): this is the CPU checker the CUDA
 * path is compared with; it is never the thing shipped or measured (except as
 * the explicitly labelled "port" CPU baseline of bench.py).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks (1) every golden vector
 * the reference's tests hold for this path (tests/golden/reference_vectors.json,
 * transcribed from test/lie_test.cpp, test/matrix_test.cpp, test/solver_test.cpp)
 * and (2) bit-for-bit equality with the reference's own headers compiled behind
 * oracle/ref_shim.cpp (-ffp-contract=off on both sides) on seeded random and
 * edge-case inputs, either live (oracle/_ref present) or through the committed
 * fixtures tests/golden/ref_outputs.npz.
 *
 * Each function follows the evaluation order of the reference routine cited
 * beside it (file:line under /root/reference), including the order in which its
 * lazy element-wise expressions and its eager k-j-i matrix product
 * (include/matrixs.hpp:596-640) accumulate, because that order is what fixes the
 * last bits.  Records are dense column-major doubles: X(r,c) = x[r + c*rows].
 */
#include <math.h>
#include <string.h>
#include <float.h>
#include <omp.h>

#include "ppx_oracle.h"

#define EPS_SP ((double)FLT_EPSILON) /* include/exprtmpl.hpp:12 */
#define EPS_DP DBL_EPSILON           /* include/exprtmpl.hpp:13 */
#define PPX_PI 3.141592653589793     /* include/exprtmpl.hpp:11 */
#define MAXD 8                       /* largest dimension handled here */

static int nthr(int nthreads) { return nthreads > 0 ? nthreads : omp_get_max_threads(); }

const char *ppxo_kind(void) { return "port"; }
int ppxo_max_threads(void) { return omp_get_max_threads(); }

/* ------------------------------------------------------------------ helpers */

/* MatrixS::operator*, include/matrixs.hpp:596-640: zero result, k-j-i accumulation. */
static void gemm(int M, int N, int L, const double *a, const double *b, double *c)
{
    for (int i = 0; i < M * L; i++)
        c[i] = 0.0;
    for (int k = 0; k < N; k++)
        for (int j = 0; j < L; j++)
            for (int i = 0; i < M; i++)
                c[i + j * M] += a[i + k * M] * b[k + j * N];
}

/* MatrixS::T(), include/matrixs.hpp:498-532 */
static void transpose(int M, int N, const double *a, double *b)
{
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++)
            b[j + i * N] = a[i + j * M];
}

static void eye(int n, double *a)
{
    memset(a, 0, sizeof(double) * n * n);
    for (int i = 0; i < n; i++)
        a[i + i * n] = 1.0;
}

/* details::is_same / near_zero, include/matrixs.hpp:34-42 */
static int is_same(double a, double b) { return fabs(a - b) < EPS_SP; }
static int near_zero(double a) { return fabs(a) < 1.0e-6; }

/* SIGN(a, b), include/linalg.hpp:26-30 */
static double sign2(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

/* norm2 of a vector = sqrt(std::inner_product(..., 0.0)), include/linalg.hpp:217-241 */
static double vnorm(const double *v, int n)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++)
        acc = acc + v[i] * v[i];
    return sqrt(acc);
}

/* hat(so3), include/liegroup.hpp:12-15 */
static void hat3(const double *w, double *m)
{
    m[0] = 0.0, m[1] = w[2], m[2] = -w[1];
    m[3] = -w[2], m[4] = 0.0, m[5] = w[0];
    m[6] = w[1], m[7] = -w[0], m[8] = 0.0;
}

/* hat(se3), include/liegroup.hpp:22-25 */
static void hat6(const double *s, double *m)
{
    m[0] = 0.0, m[1] = s[2], m[2] = -s[1], m[3] = 0.0;
    m[4] = -s[2], m[5] = 0.0, m[6] = s[0], m[7] = 0.0;
    m[8] = s[1], m[9] = -s[0], m[10] = 0.0, m[11] = 0.0;
    m[12] = s[3], m[13] = s[4], m[14] = s[5], m[15] = 0.0;
}

/* SE3::Rot / Pos, include/liegroup.hpp:209-216 */
static void rot_of(const double *T, double *R)
{
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            R[r + 3 * c] = T[r + 4 * c];
}
static void pos_of(const double *T, double *p) { p[0] = T[12], p[1] = T[13], p[2] = T[14]; }

/* SE3(const SO3&, const T3&), include/liegroup.hpp:194-198 */
static void se3_from(const double *R, const double *p, double *T)
{
    eye(4, T);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            T[r + 4 * c] = R[r + 3 * c];
    T[12] = p[0], T[13] = p[1], T[14] = p[2];
}

/* ------------------------------------------------------------- Lie algebra */

/* MatrixS<3,1>::exp, include/liegroup.hpp:116-133 */
static void so3_exp1(const double *w, double *R)
{
    double theta = vnorm(w, 3);
    if (near_zero(theta))
    {
        eye(3, R);
        return;
    }
    double omg[9], sq[9], id[9];
    hat3(w, omg);
    double coff1 = sin(theta) / theta;
    double coff2 = (1 - cos(theta)) / (theta * theta);
    gemm(3, 3, 3, omg, omg, sq);
    eye(3, id);
    for (int i = 0; i < 9; i++)
        R[i] = (id[i] + coff1 * omg[i]) + coff2 * sq[i];
}

/* SO3::log, include/liegroup.hpp:60-92 */
static void SO3_log1(const double *R, double *w)
{
    double tr = 0.0;
    for (int i = 0; i < 3; i++)
        tr += R[i + 3 * i];
    double acosinput = (tr - 1) / 2.0;
    if (acosinput >= 1.0)
    {
        w[0] = w[1] = w[2] = 0.0;
    }
    else if (acosinput <= -1.0)
    {
        double s, v[3];
        if (!is_same(1 + R[2 + 3 * 2], 0.0))
        {
            s = 1.0 / sqrt(2 + 2 * R[2 + 3 * 2]);
            v[0] = R[0 + 3 * 2], v[1] = R[1 + 3 * 2], v[2] = 1 + R[2 + 3 * 2];
        }
        else if (!is_same(1 + R[1 + 3 * 1], 0.0))
        {
            s = 1.0 / sqrt(2 + 2 * R[1 + 3 * 1]);
            v[0] = R[0 + 3 * 1], v[1] = 1 + R[1 + 3 * 1], v[2] = R[2 + 3 * 1];
        }
        else
        {
            s = 1.0 / sqrt(2 + 2 * R[0 + 3 * 0]);
            v[0] = 1 + R[0 + 3 * 0], v[1] = R[1 + 3 * 0], v[2] = R[2 + 3 * 0];
        }
        for (int i = 0; i < 3; i++)
            w[i] = PPX_PI * (s * v[i]);
    }
    else
    {
        double Rt[9], ret[9];
        double theta = acos(acosinput);
        double den = 2.0 * sin(theta);
        transpose(3, 3, R, Rt);
        for (int i = 0; i < 9; i++)
            ret[i] = (R[i] - Rt[i]) * theta / den;
        w[0] = ret[5], w[1] = ret[6], w[2] = ret[1]; /* vee, liegroup.hpp:17-20 */
    }
}

/* MatrixS<6,1>::exp, include/liegroup.hpp:253-269 */
static void se3_exp1(const double *s, double *T)
{
    double phi = vnorm(s, 3);
    if (is_same(phi, 0))
    {
        double id[9];
        eye(3, id);
        se3_from(id, s + 3, T);
        return;
    }
    double ksi[16], k2[16], k3[16], id[16];
    hat6(s, ksi);
    double coff1 = (1 - cos(phi)) / (phi * phi);
    double coff2 = (phi - sin(phi)) / (phi * phi * phi);
    gemm(4, 4, 4, ksi, ksi, k2);
    gemm(4, 4, 4, k2, ksi, k3);
    eye(4, id);
    for (int i = 0; i < 16; i++)
        T[i] = ((id[i] + ksi[i]) + coff1 * k2[i]) + coff2 * k3[i];
}

/* SE3::log, include/liegroup.hpp:231-250 */
static void SE3_log1(const double *T, double *s)
{
    double R[9], p[3];
    rot_of(T, R);
    pos_of(T, p);
    double tr = 0.0;
    for (int i = 0; i < 3; i++)
        tr += R[i + 3 * i];
    double ctheta = (tr - 1) / 2.0;
    if (fabs(ctheta) < 1)
    {
        double omg[3], om[9], sq[9], id[9], le[9], v[3];
        double theta = acos(ctheta);
        SO3_log1(R, omg);
        hat3(omg, om);
        double coff1 = (1.0 / theta - 1.0 / (tan(theta / 2.0) * 2.0)) / theta;
        gemm(3, 3, 3, om, om, sq);
        eye(3, id);
        for (int i = 0; i < 9; i++)
            le[i] = (id[i] - om[i] / 2.0) + coff1 * sq[i];
        gemm(3, 3, 1, le, p, v);
        s[0] = omg[0], s[1] = omg[1], s[2] = omg[2];
        s[3] = v[0], s[4] = v[1], s[5] = v[2];
    }
    else
    {
        /* also taken at theta = pi (ctheta <= -1): the rotation part is dropped */
        s[0] = s[1] = s[2] = 0.0;
        s[3] = p[0], s[4] = p[1], s[5] = p[2];
    }
}

/* SE3::Adt, include/liegroup.hpp:217-226 */
static void SE3_Adt1(const double *T, double *Ad)
{
    double R[9], p[3], hp[9], pr[9];
    rot_of(T, R);
    pos_of(T, p);
    hat3(p, hp);
    gemm(3, 3, 3, hp, R, pr);
    memset(Ad, 0, sizeof(double) * 36);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            Ad[r + 6 * c] = R[r + 3 * c];
            Ad[(r + 3) + 6 * c] = pr[r + 3 * c];
            Ad[(r + 3) + 6 * (c + 3)] = R[r + 3 * c];
        }
}

/* SE3::I, include/liegroup.hpp:227-230 */
static void SE3_inv1(const double *T, double *Ti)
{
    double R[9], Rt[9], p[3], q[3];
    rot_of(T, R);
    pos_of(T, p);
    transpose(3, 3, R, Rt);
    gemm(3, 3, 1, Rt, p, q);
    for (int i = 0; i < 3; i++)
        q[i] = -1 * q[i];
    se3_from(Rt, q, Ti);
}

/* MatrixS<6,1>::adt, include/liegroup.hpp:271-280 */
static void se3_adt1(const double *s, double *ad)
{
    double hw[9], hv[9];
    hat3(s, hw);
    hat3(s + 3, hv);
    memset(ad, 0, sizeof(double) * 36);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
        {
            ad[r + 6 * c] = hw[r + 3 * c];
            ad[(r + 3) + 6 * c] = hv[r + 3 * c];
            ad[(r + 3) + 6 * (c + 3)] = hw[r + 3 * c];
        }
}

/* ljac / ljacinv / rjac / rjacinv, include/liegroup.hpp:142-184 */
static void so3_jac1(int which, const double *w, double *J)
{
    double id[9], X[9], t[9], P[9], out[9];
    double x2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    int inv = which & 1;
    eye(3, id);
    hat3(w, X);
    if (x2 < EPS_DP)
    {
        for (int i = 0; i < 9; i++)
            out[i] = inv ? id[i] - 0.5 * X[i] : id[i] + 0.5 * X[i];
    }
    else
    {
        double x = sqrt(x2);
        if (!inv)
        {
            double c1 = (1 - cos(x)) / x2;
            double c2 = (x - sin(x)) / (x2 * x);
            for (int i = 0; i < 9; i++)
                t[i] = c2 * X[i];
            gemm(3, 3, 3, t, X, P);
            for (int i = 0; i < 9; i++)
                out[i] = (id[i] + c1 * X[i]) + P[i];
        }
        else
        {
            double c = 1 / x2 - (1 + cos(x)) / (2 * x * sin(x));
            for (int i = 0; i < 9; i++)
                t[i] = c * X[i];
            gemm(3, 3, 3, t, X, P);
            for (int i = 0; i < 9; i++)
                out[i] = (id[i] - 0.5 * X[i]) + P[i];
        }
    }
    if (which >= 2)
        transpose(3, 3, out, J);
    else
        memcpy(J, out, sizeof(out));
}

/* ------------------------------------------------------------- kinematics */

/* kinematics<N>::forwardSpace(q), demo/robotics.hpp:83-92 */
static void fk1(const double *screws, const double *home, int nj, const double *q, double *T)
{
    double cur[16], E[16], nxt[16], s[6];
    memcpy(cur, home, sizeof(cur));
    for (int i = nj - 1; i > -1; i--)
    {
        for (int k = 0; k < 6; k++)
            s[k] = screws[6 * i + k] * q[i];
        se3_exp1(s, E);
        gemm(4, 4, 4, E, cur, nxt);
        memcpy(cur, nxt, sizeof(cur));
    }
    memcpy(T, cur, sizeof(cur));
}

/* kinematics<N>::jacobiSpace(q), demo/robotics.hpp:94-106 */
static void jac1(const double *screws, int nj, const double *q, double *Js)
{
    double cur[16], E[16], nxt[16], Ad[36], s[6];
    memset(Js, 0, sizeof(double) * 6 * nj);
    memcpy(Js, screws, sizeof(double) * 6);
    eye(4, cur);
    for (int i = 1; i < nj; i++)
    {
        for (int k = 0; k < 6; k++)
            s[k] = screws[6 * (i - 1) + k] * q[i - 1];
        se3_exp1(s, E);
        gemm(4, 4, 4, cur, E, nxt);
        memcpy(cur, nxt, sizeof(cur));
        SE3_Adt1(cur, Ad);
        gemm(6, 6, 1, Ad, screws + 6 * i, Js + 6 * i);
    }
}

/* ------------------------------------------------------------ factorisations */

/* QR<m,n> ctor + solve + linsolve<Factorization::QR>, include/linalg.hpp:535-603, 1189-1200 */
static int qr_solve1(int m, int n, const double *A0, const double *b0, double *x)
{
    double A[MAXD * MAXD], c[MAXD], d[MAXD], b[MAXD];
    int singular = 0;
    memcpy(A, A0, sizeof(double) * m * n);
    memcpy(b, b0, sizeof(double) * m);
    int kmax = m < n ? m : n;
    for (int k = 0; k < kmax; k++)
    {
        double sum = 0.0;
        for (int i = k; i < m; i++)
            sum += A[i + k * m] * A[i + k * m];
        if (fabs(sum) < EPS_DP)
            singular = 1;
        double sigma = sign2(sqrt(sum), A[k + k * m]);
        A[k + k * m] += sigma;
        c[k] = sigma * A[k + k * m];
        d[k] = -sigma;
        for (int j = k + 1; j < n; j++)
        {
            double tau = 0.0;
            for (int i = k; i < m; i++)
                tau += A[i + k * m] * A[i + j * m];
            tau = fabs(c[k]) > EPS_DP ? tau / c[k] : 0.0;
            for (int i = k; i < m; i++)
                A[i + j * m] -= tau * A[i + k * m];
        }
    }
    if (!singular)
    {
        for (int j = 0; j < n; j++)
        {
            double sum = 0.0;
            for (int i = j; i < m; i++)
                sum += A[i + j * m] * b[i];
            sum /= c[j];
            for (int i = j; i < m; i++)
                b[i] -= sum * A[i + j * m];
        }
        for (int i = n - 1; i >= 0; i--)
        {
            double sum = 0.0;
            for (int j = i + 1; j < n; j++)
                sum += A[i + j * m] * b[j];
            b[i] = (b[i] - sum) / d[i];
        }
    }
    for (int i = 0; i < n; i++)
        x[i] = b[i];
    return singular ? 3 : 0;
}

/* LU<n>, include/linalg.hpp:448-533 */
typedef struct
{
    double A[MAXD * MAXD];
    int indx[MAXD];
    int even;
    int singular;
} lu_t;

static void lu_factor(int n, const double *A0, lu_t *f)
{
    double *A = f->A;
    memcpy(A, A0, sizeof(double) * n * n);
    f->even = 1;
    f->singular = 0;
    for (int i = 0; i < n; i++)
        f->indx[i] = i;
    for (int k = 0; k < n; k++)
    {
        double valmax = fabs(A[k + k * n]);
        int ip = k;
        for (int row = k + 1; row < n; row++)
        {
            double tmp = fabs(A[row + k * n]);
            if (valmax < tmp)
            {
                valmax = tmp;
                ip = row;
            }
        }
        if (valmax < EPS_DP)
        {
            f->singular = 1;
            break;
        }
        if (ip != k)
        {
            for (int col = 0; col < n; col++)
            {
                double t = A[ip + col * n];
                A[ip + col * n] = A[k + col * n];
                A[k + col * n] = t;
            }
            int ti = f->indx[ip];
            f->indx[ip] = f->indx[k];
            f->indx[k] = ti;
            f->even = !f->even;
        }
        for (int row = k + 1; row < n; row++)
        {
            double weight = A[row + k * n] / A[k + k * n];
            A[row + k * n] = weight;
            for (int col = k + 1; col < n; col++)
                A[row + col * n] -= weight * A[k + col * n];
        }
    }
}

static void lu_solve_vec(int n, const lu_t *f, double *b)
{
    const double *A = f->A;
    double y[MAXD];
    y[0] = b[f->indx[0]];
    for (int row = 1; row < n; row++)
    {
        double sum = 0.0;
        for (int col = 0; col < row; col++)
            sum += A[row + col * n] * y[col];
        y[row] = b[f->indx[row]] - sum;
    }
    b[n - 1] = y[n - 1] / A[(n - 1) + (n - 1) * n];
    for (int row = n - 2; row >= 0; row--)
    {
        double sum = 0.0;
        for (int col = row + 1; col < n; col++)
            sum += A[row + col * n] * b[col];
        b[row] = (y[row] - sum) / A[row + row * n];
    }
}

/* SVD<m,n> ctor (Golub-Kahan-Reinsch as the reference states it), include/linalg.hpp:609-884.
 * u: m x n, w: n, v: n x n.  Indices are spelled U(r,c) = u[r + c*m], V(r,c) = v[r + c*n]. */
#define U(r, c) u[(r) + (c) * m]
#define V(r, c) v[(r) + (c) * n]
static void svd1(int m, int n, const double *A0, double *u, double *w, double *v)
{
    int i, its, j, jj, k, nm = 0, l = 0;
    double c, f, h, s, x, y, z;
    double rv1[MAXD];
    double g = 0.0, scale = 0.0, anorm = 0.0;
    memcpy(u, A0, sizeof(double) * m * n);
    memset(v, 0, sizeof(double) * n * n);
    memset(w, 0, sizeof(double) * n);
    memset(rv1, 0, sizeof(rv1));
    /* Householder reduction to bidiagonal form, :619-702 */
    for (i = 0; i < n; i++)
    {
        l = i + 2;
        rv1[i] = scale * g;
        g = s = scale = 0.0;
        if (i < m)
        {
            for (k = i; k < m; k++)
                scale += fabs(U(k, i));
            if (scale != 0.0)
            {
                for (k = i; k < m; k++)
                {
                    U(k, i) /= scale;
                    s += U(k, i) * U(k, i);
                }
                f = U(i, i);
                g = -sign2(sqrt(s), f);
                h = f * g - s;
                U(i, i) = f - g;
                for (j = l - 1; j < n; j++)
                {
                    for (s = 0.0, k = i; k < m; k++)
                        s += U(k, i) * U(k, j);
                    f = s / h;
                    for (k = i; k < m; k++)
                        U(k, j) += f * U(k, i);
                }
                for (k = i; k < m; k++)
                    U(k, i) *= scale;
            }
        }
        w[i] = scale * g;
        g = s = scale = 0.0;
        if (i + 1 <= m && i + 1 != n)
        {
            for (k = l - 1; k < n; k++)
                scale += fabs(U(i, k));
            if (scale != 0.0)
            {
                for (k = l - 1; k < n; k++)
                {
                    U(i, k) /= scale;
                    s += U(i, k) * U(i, k);
                }
                f = U(i, l - 1);
                g = -sign2(sqrt(s), f);
                h = f * g - s;
                U(i, l - 1) = f - g;
                for (k = l - 1; k < n; k++)
                    rv1[k] = U(i, k) / h;
                for (j = l - 1; j < m; j++)
                {
                    for (s = 0.0, k = l - 1; k < n; k++)
                        s += U(j, k) * U(i, k);
                    for (k = l - 1; k < n; k++)
                        U(j, k) += s * rv1[k];
                }
                for (k = l - 1; k < n; k++)
                    U(i, k) *= scale;
            }
        }
        anorm = fmax(anorm, fabs(w[i]) + fabs(rv1[i]));
    }
    /* accumulate right-hand transformations, :703-734 */
    for (i = n - 1; i >= 0; i--)
    {
        if (i < n - 1)
        {
            if (g != 0.0)
            {
                for (j = l; j < n; j++)
                    V(j, i) = (U(i, j) / U(i, l)) / g;
                for (j = l; j < n; j++)
                {
                    for (s = 0.0, k = l; k < n; k++)
                        s += U(i, k) * V(k, j);
                    for (k = l; k < n; k++)
                        V(k, j) += s * V(k, i);
                }
            }
            for (j = l; j < n; j++)
                V(i, j) = V(j, i) = 0.0;
        }
        V(i, i) = 1.0;
        g = rv1[i];
        l = i;
    }
    /* accumulate left-hand transformations, :735-771 */
    for (i = (m < n ? m : n) - 1; i >= 0; i--)
    {
        l = i + 1;
        g = w[i];
        for (j = l; j < n; j++)
            U(i, j) = 0.0;
        if (g != 0.0)
        {
            g = 1.0 / g;
            for (j = l; j < n; j++)
            {
                for (s = 0.0, k = l; k < m; k++)
                    s += U(k, i) * U(k, j);
                f = (s / U(i, i)) * g;
                for (k = i; k < m; k++)
                    U(k, j) += f * U(k, i);
            }
            for (j = i; j < m; j++)
                U(j, i) *= g;
        }
        else
        {
            for (j = i; j < m; j++)
                U(j, i) = 0.0;
        }
        U(i, i) += 1.0;
    }
    /* diagonalisation of the bidiagonal form, :772-883 */
    for (k = n - 1; k >= 0; k--)
    {
        for (its = 0; its < 30; its++)
        {
            int flag = 1;
            for (l = k; l >= 0; l--)
            {
                nm = l - 1;
                if (l == 0 || fabs(rv1[l]) < EPS_DP * anorm)
                {
                    flag = 0;
                    break;
                }
                if (fabs(w[nm]) < EPS_DP * anorm)
                    break;
            }
            if (flag)
            {
                c = 0.0;
                s = 1.0;
                for (i = l; i < k + 1; i++)
                {
                    f = s * rv1[i];
                    rv1[i] = c * rv1[i];
                    if (fabs(f) < EPS_DP * anorm)
                        break;
                    g = w[i];
                    h = sqrt(f * f + g * g);
                    w[i] = h;
                    h = 1.0 / h;
                    c = g * h;
                    s = -f * h;
                    for (j = 0; j < m; j++)
                    {
                        y = U(j, nm);
                        z = U(j, i);
                        U(j, nm) = y * c + z * s;
                        U(j, i) = z * c - y * s;
                    }
                }
            }
            z = w[k];
            if (l == k)
            {
                if (z < 0.0)
                {
                    w[k] = -z;
                    for (j = 0; j < n; j++)
                        V(j, k) *= -1;
                }
                break;
            }
            x = w[l];
            nm = k - 1;
            y = w[nm];
            g = rv1[nm];
            h = rv1[k];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
            g = sqrt(f * f + 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + sign2(g, f))) - h)) / x;
            c = s = 1.0;
            for (j = l; j <= nm; j++)
            {
                i = j + 1;
                g = rv1[i];
                y = w[i];
                h = s * g;
                g = c * g;
                z = sqrt(f * f + h * h);
                rv1[j] = z;
                c = f / z;
                s = h / z;
                f = x * c + g * s;
                g = g * c - x * s;
                h = y * s;
                y *= c;
                for (jj = 0; jj < n; jj++)
                {
                    x = V(jj, j);
                    z = V(jj, i);
                    V(jj, j) = x * c + z * s;
                    V(jj, i) = z * c - x * s;
                }
                z = sqrt(f * f + h * h);
                w[j] = z;
                if (fabs(z) > EPS_DP)
                {
                    z = 1.0 / z;
                    c = f * z;
                    s = h * z;
                }
                f = c * g + s * y;
                x = c * y - s * g;
                for (jj = 0; jj < m; jj++)
                {
                    y = U(jj, j);
                    z = U(jj, i);
                    U(jj, j) = y * c + z * s;
                    U(jj, i) = z * c - y * s;
                }
            }
            rv1[l] = 0.0;
            rv1[k] = f;
            w[k] = x;
        }
    }
}

static double wmax(const double *w, int n)
{
    double mx = w[0];
    for (int i = 1; i < n; i++)
        if (mx < w[i])
            mx = w[i];
    return mx;
}

/* SVD::solve, include/linalg.hpp:886-913 (b has max(m, n) slots; the first n are the answer) */
static void svd_backsub(int m, int n, const double *u, const double *w, const double *v, double *b)
{
    double tmp[MAXD];
    double tsh = 0.5 * sqrt((double)(m + n + 1)) * wmax(w, n) * EPS_DP;
    for (int j = 0; j < n; j++)
    {
        double s = 0.0;
        if (w[j] > tsh)
        {
            for (int i = 0; i < m; i++)
                s += U(i, j) * b[i];
            s /= w[j];
        }
        tmp[j] = s;
    }
    for (int j = 0; j < n; j++)
    {
        double s = 0.0;
        for (int jj = 0; jj < n; jj++)
            s += V(j, jj) * tmp[jj];
        b[j] = s;
    }
}
#undef U
#undef V

/* EigenValue<n>: tred2 + tliq + eigsrt, include/linalg.hpp:967-1167 */
#define Z(r, c) z[(r) + (c) * n]
static void eig1(int n, const double *S, int sorted, double *d, double *z)
{
    double e[MAXD];
    memcpy(z, S, sizeof(double) * n * n);
    memset(d, 0, sizeof(double) * n);
    memset(e, 0, sizeof(e));
    /* tred2, :973-1063 */
    for (int i = n - 1; i > 0; i--)
    {
        int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0)
        {
            for (int k = 0; k < i; k++)
                scale += fabs(Z(i, k));
            if (fabs(scale) < EPS_DP)
            {
                e[i] = Z(i, l);
            }
            else
            {
                for (int k = 0; k < i; k++)
                {
                    Z(i, k) /= scale;
                    h += Z(i, k) * Z(i, k);
                }
                double f = Z(i, l);
                double g = f > 0.0 ? -sqrt(h) : sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                Z(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j < i; j++)
                {
                    Z(j, i) = Z(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k < j + 1; k++)
                        g += Z(j, k) * Z(i, k);
                    for (int k = j + 1; k < i; k++)
                        g += Z(k, j) * Z(i, k);
                    e[j] = g / h;
                    f += e[j] * Z(i, j);
                }
                double hh = f / (2 * h);
                for (int j = 0; j < i; j++)
                {
                    f = Z(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k < j + 1; k++)
                        Z(j, k) -= f * e[k] + g * Z(i, k);
                }
            }
        }
        else
        {
            e[i] = Z(i, l);
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; i++)
    {
        if (fabs(d[i]) > EPS_DP)
        {
            for (int j = 0; j < i; j++)
            {
                double g = 0.0;
                for (int k = 0; k < i; k++)
                    g += Z(i, k) * Z(k, j);
                for (int k = 0; k < i; k++)
                    Z(k, j) -= g * Z(k, i);
            }
        }
        d[i] = Z(i, i);
        Z(i, i) = 1.0;
        for (int j = 0; j < i; j++)
            Z(j, i) = Z(i, j) = 0.0;
    }
    /* tliq (implicit QL), :1065-1135 */
    for (int i = 1; i < n; i++)
        e[i - 1] = e[i];
    e[n - 1] = 0.0;
    int give_up = 0;
    for (int l = 0; l < n && !give_up; l++)
    {
        int iter = 0, m;
        do
        {
            for (m = l; m < n - 1; m++)
                if (fabs(e[m]) < EPS_DP * (fabs(d[m]) + fabs(d[m + 1])))
                    break;
            if (m != l)
            {
                if (iter++ == 30)
                {
                    give_up = 1; /* the reference returns silently */
                    break;
                }
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(g * g + 1.0);
                g = d[m] - d[l] + e[l] / (g + sign2(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--)
                {
                    double f = s * e[i];
                    double b = c * e[i];
                    r = sqrt(f * f + g * g);
                    e[i + 1] = r;
                    if (fabs(r) < EPS_DP)
                    {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int k = 0; k < n; k++)
                    {
                        f = Z(k, i + 1);
                        Z(k, i + 1) = s * Z(k, i) + c * f;
                        Z(k, i) = c * Z(k, i) - s * f;
                    }
                }
                if (fabs(r) < EPS_DP && i >= l)
                    continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    /* eigsrt (ascending selection sort), :1137-1151 */
    if (sorted)
    {
        for (int i = 0; i < n - 1; i++)
        {
            int j = i;
            for (int t = i + 1; t < n; t++)
                if (d[t] < d[j])
                    j = t;
            if (j != i)
            {
                double t = d[i];
                d[i] = d[j];
                d[j] = t;
                for (int k = 0; k < n; k++)
                {
                    t = Z(k, i);
                    Z(k, i) = Z(k, j);
                    Z(k, j) = t;
                }
            }
        }
    }
}
#undef Z

/* one Newton step, test/lie_test.cpp:104-110, on demo/robotics.hpp's kinematics */
static void ik_step1(const double *screws, const double *home, int nj, const double *q, const double *Tt,
                     double *q_out, double *Vs)
{
    double Tsb[16], Ad[36], Ti[16], X[16], lg[6], Js[6 * MAXD];
    double u[6 * MAXD], w[MAXD], v[MAXD * MAXD], b[MAXD];
    fk1(screws, home, nj, q, Tsb);
    SE3_Adt1(Tsb, Ad);
    SE3_inv1(Tsb, Ti);
    gemm(4, 4, 4, Ti, Tt, X);
    SE3_log1(X, lg);
    gemm(6, 6, 1, Ad, lg, Vs);
    jac1(screws, nj, q, Js);
    svd1(6, nj, Js, u, w, v);
    memcpy(b, Vs, sizeof(double) * 6);
    svd_backsub(6, nj, u, w, v, b);
    for (int i = 0; i < nj; i++)
        q_out[i] = q[i] + b[i];
}

/* ------------------------------------------------------------ batch loops */

#define LOOP(...)                                                                   \
    {                                                                               \
        _Pragma("omp parallel for schedule(static) num_threads(nthr(nthreads))")    \
        for (long long i = 0; i < (long long)count; i++) { __VA_ARGS__ }            \
        return 0;                                                                   \
    }

int ppxo_so3_exp(const double *w, double *R, size_t count, int nthreads) LOOP(so3_exp1(w + 3 * i, R + 9 * i);)
int ppxo_SO3_log(const double *R, double *w, size_t count, int nthreads) LOOP(SO3_log1(R + 9 * i, w + 3 * i);)
int ppxo_se3_exp(const double *xi, double *T, size_t count, int nthreads) LOOP(se3_exp1(xi + 6 * i, T + 16 * i);)
int ppxo_SE3_log(const double *T, double *xi, size_t count, int nthreads) LOOP(SE3_log1(T + 16 * i, xi + 6 * i);)
int ppxo_se3_exp_log(const double *xi, double *xo, size_t count, int nthreads)
    LOOP(double T[16]; se3_exp1(xi + 6 * i, T); SE3_log1(T, xo + 6 * i);)
/* MatrixS::operator*, include/matrixs.hpp:596-640 */
int ppxo_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count, int nthreads)
{
    if (m < 1 || n < 1 || l < 1)
        return -1;
    LOOP(gemm(m, n, l, A + (size_t)m * n * i, B + (size_t)n * l * i, C + (size_t)m * l * i);)
}

int ppxo_SE3_mul(const double *A, const double *B, double *C, size_t count, int nthreads)
    LOOP(gemm(4, 4, 4, A + 16 * i, B + 16 * i, C + 16 * i);)
int ppxo_SE3_inv(const double *T, double *Ti, size_t count, int nthreads) LOOP(SE3_inv1(T + 16 * i, Ti + 16 * i);)
int ppxo_SE3_Adt(const double *T, double *Ad, size_t count, int nthreads) LOOP(SE3_Adt1(T + 16 * i, Ad + 36 * i);)
int ppxo_se3_adt(const double *xi, double *ad, size_t count, int nthreads) LOOP(se3_adt1(xi + 6 * i, ad + 36 * i);)

int ppxo_so3_jac(int which, const double *w, double *J, size_t count, int nthreads)
{
    if (which < 0 || which > 3)
        return -1;
    LOOP(so3_jac1(which, w + 3 * i, J + 9 * i);)
}

int ppxo_poe_fk_jacobian(const double *screws, const double *home, int nj, const double *q, double *T, double *Js,
                         size_t count, int nthreads)
{
    if (nj < 1 || nj > MAXD)
        return -1;
    LOOP(if (T) fk1(screws, home, nj, q + nj * i, T + 16 * i);
         if (Js) jac1(screws, nj, q + nj * i, Js + 6 * nj * i);)
}

int ppxo_linsolve_qr(int m, int n, const double *A, const double *b, double *x, signed char *status, size_t count,
                     int nthreads)
{
    if (m < n || n < 1 || m > MAXD)
        return -1;
    LOOP(int s = qr_solve1(m, n, A + m * n * i, b + m * i, x + n * i); if (status) status[i] = (signed char)s;)
}

int ppxo_linsolve_lu(int n, const double *A, const double *b, double *x, signed char *status, size_t count,
                     int nthreads)
{
    if (n < 1 || n > MAXD)
        return -1;
    LOOP(lu_t f; lu_factor(n, A + n * n * i, &f); memcpy(x + n * i, b + n * i, sizeof(double) * n);
         if (!f.singular) lu_solve_vec(n, &f, x + n * i); if (status) status[i] = f.singular ? 3 : 0;)
}

int ppxo_inverse(int n, const double *A, double *Ai, size_t count, int nthreads)
{
    if (n < 1 || n > MAXD)
        return -1;
    /* MatrixS::I(), include/linalg.hpp:933-948 */
    LOOP(lu_t f; lu_factor(n, A + n * n * i, &f); double *out = Ai + n * n * i;
         if (f.singular) { memset(out, 0, sizeof(double) * n * n); } else {
             eye(n, out);
             for (int j = 0; j < n; j++) lu_solve_vec(n, &f, out + j * n);
         })
}

int ppxo_det(int n, const double *A, double *det, size_t count, int nthreads)
{
    if (n < 1 || n > MAXD)
        return -1;
    /* MatrixS::det(), include/linalg.hpp:950-965 */
    LOOP(lu_t f; lu_factor(n, A + n * n * i, &f); double D = 0.0; if (!f.singular) {
        D = f.even ? 1.0 : -1.0;
        for (int j = 0; j < n; j++) D *= f.A[j + j * n];
    } det[i] = D;)
}

int ppxo_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count, int nthreads)
{
    if (m < 1 || n < 1 || m > MAXD || n > MAXD)
        return -1;
    LOOP(double u[MAXD * MAXD]; double v[MAXD * MAXD]; svd1(m, n, A + m * n * i, u, w + n * i, v);
         if (U) memcpy(U + m * n * i, u, sizeof(double) * m * n);
         if (V) memcpy(V + n * n * i, v, sizeof(double) * n * n);)
}

int ppxo_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count, int nthreads)
{
    if (m < n || n < 1 || m > MAXD) /* linsolve<SVD> returns b.sub<N,1>: needs n <= m */
        return -1;
    LOOP(double u[MAXD * MAXD]; double v[MAXD * MAXD]; double w[MAXD]; double bb[MAXD];
         svd1(m, n, A + m * n * i, u, w, v); memcpy(bb, b + m * i, sizeof(double) * m);
         svd_backsub(m, n, u, w, v, bb); memcpy(x + n * i, bb, sizeof(double) * n);)
}

int ppxo_pinv(int m, int n, const double *A, double *P, size_t count, int nthreads)
{
    if (m < 1 || n < 1 || m > MAXD || n > MAXD)
        return -1;
    /* pinv, include/linalg.hpp:1211-1227: (v * W) * u^T with eager products */
    LOOP(double u[MAXD * MAXD]; double v[MAXD * MAXD]; double w[MAXD]; double W[MAXD * MAXD]; double vW[MAXD * MAXD];
         double ut[MAXD * MAXD];
         svd1(m, n, A + m * n * i, u, w, v);
         double tsh = 0.5 * sqrt((double)(m + n + 1)) * wmax(w, n) * EPS_DP;
         memset(W, 0, sizeof(double) * n * n);
         for (int j = 0; j < n; j++) if (fabs(w[j]) > tsh) W[j + j * n] = 1.0 / w[j];
         gemm(n, n, n, v, W, vW); transpose(m, n, u, ut); gemm(n, n, m, vW, ut, P + m * n * i);)
}

int ppxo_norm2(int m, int n, const double *A, double *s, size_t count, int nthreads)
{
    if (m < 2 || n < 2 || m > MAXD || n > MAXD)
        return -1;
    LOOP(double u[MAXD * MAXD]; double v[MAXD * MAXD]; double w[MAXD]; svd1(m, n, A + m * n * i, u, w, v);
         s[i] = wmax(w, n);)
}

int ppxo_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count, int nthreads)
{
    if (n < 1 || n > MAXD)
        return -1;
    LOOP(double zz[MAXD * MAXD]; eig1(n, S + n * n * i, sorted, d + n * i, zz);
         if (z) memcpy(z + n * n * i, zz, sizeof(double) * n * n);)
}

int ppxo_ik_newton_step(const double *screws, const double *home, int nj, const double *q, const double *Tt,
                        double *q_out, double *Vs, size_t count, int nthreads)
{
    if (nj < 1 || nj > 6)
        return -1;
    LOOP(double vs[6]; ik_step1(screws, home, nj, q + nj * i, Tt + 16 * i, q_out + nj * i, vs);
         if (Vs) memcpy(Vs + 6 * i, vs, sizeof(vs));)
}

int ppxo_ik_solve(const double *screws, const double *home, int nj, const double *q0, const double *Tt,
                  int max_iter, double *q_out, int *iters, size_t count, int nthreads)
{
    if (nj < 1 || nj > 6)
        return -1;
    /* test/lie_test.cpp:96-113 */
    LOOP(double q[MAXD]; double qn[MAXD]; double vs[6]; memcpy(q, q0 + nj * i, sizeof(double) * nj);
         int conv = 0; int iter = 0;
         while (!conv && iter < max_iter) {
             ik_step1(screws, home, nj, q, Tt + 16 * i, qn, vs);
             conv = vnorm(vs, 3) < EPS_SP && vnorm(vs + 3, 3) < EPS_SP;
             memcpy(q, qn, sizeof(double) * nj);
             ++iter;
         }
         memcpy(q_out + nj * i, q, sizeof(double) * nj); if (iters) iters[i] = iter;)
}

/* ------------------------------------------------------------ bounded trust-region IK (CoDo) */

#define MAX_SP ((double)FLT_MAX) /* include/exprtmpl.hpp:14 */
static double sqr(double a) { return a * a; }

/* the residual and Jacobian kinematics<N>::inverseSpace hands to CoDo, demo/robotics.hpp:128-140 */
static _Thread_local long g_codo_fevals; /* diagnostic: residual evaluations of the current codo6 call */
static void codo_f(const double *screws, const double *home, int nj, const double *poseI, const double *q, double *fx)
{
    g_codo_fevals++;
    double Tsb[16], X[16];
    fk1(screws, home, nj, q, Tsb);
    gemm(4, 4, 4, Tsb, poseI, X);
    SE3_log1(X, fx);
}
static void codo_df(const double *screws, const double *home, int nj, const double *poseI, const double *q,
                    double *jac)
{
    double Tsb[16], X[16], p[3], hp[9], JpI[36], Js[6 * MAXD];
    fk1(screws, home, nj, q, Tsb);
    gemm(4, 4, 4, Tsb, poseI, X);
    pos_of(X, p);
    hat3(p, hp);
    eye(6, JpI);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            JpI[(r + 3) + 6 * c] = -1 * hp[r + 3 * c];
    jac1(screws, nj, q, Js);
    gemm(6, 6, nj, JpI, Js, jac);
}

static double dot_n(const double *a, const double *b, int n)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++)
        acc = acc + a[i] * b[i];
    return acc;
}
static double vmin(const double *a, int n)
{
    double m = a[0];
    for (int i = 1; i < n; i++)
        if (a[i] < m)
            m = a[i];
    return m;
}

/* details::CoDo<6,6>::operator(), include/optimization.hpp:542-779, with DMatrixCi (:787-808).
 * Returns the status (1 CONVERGED / 2 DIVERGED); x is updated in place, *y = 0.5 |f|. */
static int codo6(const double *screws, const double *home, const double *poseI, const double *lo, const double *hi,
                 double *x, double *y)
{
    enum { N = 6 };
    const double sigma = 0.99995, theta = 0.99995, FTOLA = EPS_SP;
    const unsigned ITMAX = 300;
    for (int i = 0; i < N; i++)
        if (!(lo[i] <= x[i] && x[i] <= hi[i]))
        {
            *y = MAX_SP;
            return 2;
        }
    double fx[N], grad[N] = {0}, p[N] = {0};
    codo_f(screws, home, N, poseI, x, fx);
    double fnrm = vnorm(fx, N);
    unsigned itc = 0;
    double delta0 = 1.0;
    int stat = 1;
    while (fnrm > FTOLA && itc++ < ITMAX)
    {
        double fnrm_old = fnrm, grad_old[N], jac[36], jacT[36], tmp[N], d[N], dsqrt[N], G[N], dgrad[N], pc[N], pn[N];
        memcpy(grad_old, grad, sizeof(grad));
        codo_df(screws, home, N, poseI, x, jac);
        transpose(6, 6, jac, jacT);
        gemm(6, 6, 1, jacT, fx, grad);
        double lambda;
        if (itc == 1)
            lambda = fmax(1e-2, vnorm(grad, N));
        else
        {
            for (int i = 0; i < N; i++)
                tmp[i] = grad[i] - grad_old[i];
            lambda = fmax(1e-2, dot_n(p, tmp, N) / dot_n(p, p, N));
        }
        for (int i = 0; i < N; i++) /* Hager scaling */
        {
            if (grad[i] < 0.0 && hi[i] <= MAX_SP)
                d[i] = 1.0 / (lambda - grad[i] / (hi[i] - x[i]));
            else if (grad[i] > 0 && lo[i] >= -MAX_SP)
                d[i] = 1.0 / (lambda + grad[i] / (x[i] - lo[i]));
            else
                d[i] = 1.0 / lambda;
        }
        if (vmin(d, N) < EPS_DP)
        {
            stat = 2;
            break;
        }
        for (int i = 0; i < N; i++)
        {
            dsqrt[i] = sqrt(d[i]);
            G[i] = 1.0 / dsqrt[i];
            dgrad[i] = d[i] * grad[i];
        }
        for (int i = 0; i < N; i++)
            tmp[i] = G[i] * dgrad[i];
        double nGdgrad = vnorm(tmp, N);
        if (vnorm(dgrad, N) < FTOLA)
            break;
        double jd[N];
        for (int i = 0; i < N; i++)
            tmp[i] = dsqrt[i] * grad[i];
        gemm(6, 6, 1, jac, dgrad, jd);
        double vert = sqr(vnorm(tmp, N) / vnorm(jd, N));
        for (int i = 0; i < N; i++)
            pc[i] = -vert * dgrad[i];
        if (itc == 1)
        {
            for (int i = 0; i < N; i++)
                tmp[i] = grad[i] / G[i];
            delta0 = vnorm(tmp, N);
        }
        lu_t f;
        double rx[N];
        lu_factor(N, jac, &f);
        memcpy(rx, fx, sizeof(rx));
        if (!f.singular)
            lu_solve_vec(N, &f, rx);
        for (int i = 0; i < N; i++)
            pn[i] = -1 * rx[i];
        if (!f.singular)
        {
            for (int i = 0; i < N; i++)
                pn[i] = fmax(x[i] + pn[i], lo[i]) - x[i];
            for (int i = 0; i < N; i++)
                pn[i] = fmin(x[i] + pn[i], hi[i]) - x[i];
            double sc = fmax(sigma, 1 - vnorm(pn, N));
            for (int i = 0; i < N; i++)
                pn[i] = sc * pn[i];
        }
        double rho = 0.0, delta1 = delta0, xp[N], fxp[N], fxpnrm;
        int again;
        do
        {
            double pciv[N], alp[N];
            memcpy(pciv, pc, sizeof(pciv));
            if (vert * nGdgrad > delta1)
                for (int i = 0; i < N; i++)
                    pciv[i] = -delta1 / nGdgrad * dgrad[i];
            for (int i = 0; i < N; i++)
                alp[i] = fabs(pciv[i]) < EPS_DP ? MAX_SP : fmax((lo[i] - x[i]) / pciv[i], (hi[i] - x[i]) / pciv[i]);
            double alpha1 = vmin(alp, N);
            if (alpha1 <= 1)
            {
                double sc = fmax(theta, 1 - vnorm(pciv, N)) * alpha1;
                for (int i = 0; i < N; i++)
                    pciv[i] *= sc;
            }
            if (!f.singular)
            {
                double aa[N], seg[N], bb[N], jp[N];
                gemm(6, 6, 1, jac, pciv, jp);
                for (int i = 0; i < N; i++)
                {
                    aa[i] = fx[i] + jp[i];
                    seg[i] = pn[i] - pciv[i];
                }
                gemm(6, 6, 1, jac, seg, bb);
                double gamma = 0.0;
                if (vnorm(bb, N) != 0)
                {
                    double gamma_min = -dot_n(aa, bb, N) / dot_n(bb, bb, N);
                    double Gseg[N], Gpciv[N];
                    for (int i = 0; i < N; i++)
                    {
                        Gseg[i] = G[i] * seg[i];
                        Gpciv[i] = G[i] * pciv[i];
                    }
                    double a = sqr(vnorm(Gseg, N));
                    double b = dot_n(Gpciv, Gseg, N);
                    double c = sqr(vnorm(Gpciv, N)) - sqr(delta1);
                    double l1 = (-b + sign2(sqrt(b * b - a * c), -b)) / a;
                    double l2 = c / (l1 * a);
                    if (gamma_min > 0)
                    {
                        double gamma_end = fmax(l1, l2);
                        gamma = fmin(gamma_min, gamma_end);
                        if (gamma > 1)
                        {
                            for (int i = 0; i < N; i++)
                                alp[i] = seg[i] != 0 ? fmax((lo[i] - x[i] - pciv[i]) / seg[i], (hi[i] - x[i] - pciv[i]) / seg[i])
                                                     : MAX_SP;
                            alpha1 = vmin(alp, N);
                            gamma = fmin(theta * alpha1, gamma);
                        }
                    }
                    else
                    {
                        double gamma_end = fmin(l1, l2);
                        gamma = fmax(gamma_min, gamma_end);
                        if (gamma < 0)
                        {
                            for (int i = 0; i < N; i++)
                                alp[i] = seg[i] != 0 ? fmax((x[i] + pciv[i] - lo[i]) / seg[i], (x[i] + pciv[i] - hi[i]) / seg[i])
                                                     : MAX_SP;
                            alpha1 = vmin(alp, N);
                            gamma = fmax(-theta * alpha1, gamma);
                        }
                    }
                }
                for (int i = 0; i < N; i++)
                    p[i] = (1 - gamma) * pciv[i] + gamma * pn[i];
            }
            else
                memcpy(p, pciv, sizeof(pciv));
            double jp2[N], lin[N];
            for (int i = 0; i < N; i++)
                xp[i] = x[i] + p[i];
            codo_f(screws, home, N, poseI, xp, fxp);
            fxpnrm = vnorm(fxp, N);
            gemm(6, 6, 1, jac, p, jp2);
            for (int i = 0; i < N; i++)
                lin[i] = fx[i] + jp2[i];
            rho = (fnrm - fxpnrm) / (fnrm - vnorm(lin, N));
            again = 0;
            if (rho < 0.25)
            {
                double Gp[N];
                for (int i = 0; i < N; i++)
                    Gp[i] = G[i] * p[i];
                delta1 = fmin(0.25 * delta0, 0.5 * vnorm(Gp, N));
                again = delta1 > sqrt(EPS_DP);
            }
        } while (again);
        if (rho < 0.25 && delta1 < sqrt(EPS_DP))
        {
            stat = 2;
            break;
        }
        if (rho > 0.75)
        {
            double Gp[N];
            for (int i = 0; i < N; i++)
                Gp[i] = G[i] * p[i];
            delta0 = fmax(delta1, 2 * vnorm(Gp, N));
        }
        memcpy(x, xp, sizeof(xp));
        int need_update = 0;
        for (int i = 0; i < N; i++)
        {
            if (x[i] < lo[i] + 1e3 * EPS_DP)
            {
                x[i] = lo[i] + 1e3 * EPS_DP;
                need_update = 1;
            }
            if (x[i] > hi[i] - 1e3 * EPS_DP)
            {
                x[i] = hi[i] - 1e3 * EPS_DP;
                need_update = 1;
            }
        }
        if (need_update)
        {
            codo_f(screws, home, N, poseI, x, fx);
            fnrm = vnorm(fx, N);
        }
        else
        {
            memcpy(fx, fxp, sizeof(fxp));
            fnrm = fxpnrm;
        }
        if (fabs(fnrm - fnrm_old) < FTOLA * fnrm && fnrm > FTOLA)
            break;
    }
    *y = 0.5 * fnrm;
    return stat;
}

int ppxo_ik_codo(const double *screws, const double *home, int nj, const double *lo, const double *hi,
                 const double *Tt, const double *init, double *q_out, double *x_raw, double *fnorm,
                 signed char *status, size_t count, int nthreads)
{
    if (nj != 6)
        return -1;
    LOOP(double poseI[16]; double x[6]; double y; SE3_inv1(Tt + 16 * i, poseI); memcpy(x, init + 6 * i, sizeof(x));
         int st = codo6(screws, home, poseI, lo, hi, x, &y);
         /* inverseSpace: result.s == CONVERGED ? result.x : init (demo/robotics.hpp:151) */
         memcpy(q_out + 6 * i, st == 1 ? x : init + 6 * i, sizeof(x));
         if (x_raw) memcpy(x_raw + 6 * i, x, sizeof(x)); if (fnorm) fnorm[i] = y; if (status) status[i] = (signed char)st;)
}

/* port-only diagnostic (not part of ppx_oracle.h): residual evaluations per instance of ppxo_ik_codo */
int ppxo_ik_codo_fevals(const double *screws, const double *home, const double *lo, const double *hi, const double *Tt,
                        const double *init, long *fevals, size_t count, int nthreads)
{
    LOOP(double poseI[16]; double x[6]; double y; SE3_inv1(Tt + 16 * i, poseI); memcpy(x, init + 6 * i, sizeof(x));
         g_codo_fevals = 0; codo6(screws, home, poseI, lo, hi, x, &y); fevals[i] = g_codo_fevals;)
}
