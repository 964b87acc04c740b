/*
 * ppx_oracle.h -- C interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (matrix_b200/, include/)
 * may include, link or dlopen anything from this directory; only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py do, and there only as the checker or the timed CPU baseline.
 *
 * Two libraries export exactly this interface:
 *   oracle/_ref/libppx_ref.so      ref_shim.cpp: thin extern "C" loops over the
 *                                  UNMODIFIED reference headers, compiled from
 *                                  where they lie under /root/reference (kind
 *                                  "reference").
 *   oracle/libppx_oracle.so        ppx_oracle.c: a plain-C restatement of the
 *                                  same algorithms (kind "port"), pinned
 *                                  bit-for-bit to the one above by
 *                                  tests/test_oracle.py and to the reference's
 *                                  golden vectors by tests/golden/.
 *
 * Data layout: arrays of dense reference records ("AoS"), i.e. exactly
 * std::vector<ppx::MatrixS<M,N>>::data() -- column-major doubles, M*N per
 * instance, no padding (reference include/matrixs.hpp:52-63, 572-582).
 *
 * Every function loops over `count` independent instances with an OpenMP
 * `parallel for schedule(static)`; `nthreads` <= 0 means "all host cores".
 * Return value: 0, or -1 for a shape the library was not instantiated for.
 */
#ifndef PPX_ORACLE_H
#define PPX_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *ppxo_kind(void);    /* "reference" | "port" */
int ppxo_max_threads(void);

/* MatrixS<m,n>::operator*(MatrixS<n,l>), matrixs.hpp:596-640; the reference shim instantiates
 * (3,3,3) (4,4,4) (6,6,6) (6,6,1) (4,4,1) (3,3,1) (4,3,3) (5,4,2) (2,7,3), the port takes any shape */
int ppxo_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count, int nthreads);
/* liegroup.hpp:116-133 / :60-92 */
int ppxo_so3_exp(const double *w, double *R, size_t count, int nthreads);
int ppxo_SO3_log(const double *R, double *w, size_t count, int nthreads);
/* liegroup.hpp:253-269 / :231-250 */
int ppxo_se3_exp(const double *xi, double *T, size_t count, int nthreads);
int ppxo_SE3_log(const double *T, double *xi, size_t count, int nthreads);
int ppxo_se3_exp_log(const double *xi, double *xi_out, size_t count, int nthreads);
/* liegroup.hpp:204-207, :227-230, :217-226, :271-280 */
int ppxo_SE3_mul(const double *A, const double *B, double *C, size_t count, int nthreads);
int ppxo_SE3_inv(const double *T, double *Ti, size_t count, int nthreads);
int ppxo_SE3_Adt(const double *T, double *Ad, size_t count, int nthreads);
int ppxo_se3_adt(const double *xi, double *ad, size_t count, int nthreads);
/* so3 Jacobian family, liegroup.hpp:142-184; which: 0 ljac 1 ljacinv 2 rjac 3 rjacinv */
int ppxo_so3_jac(int which, const double *w, double *J, size_t count, int nthreads);

/* demo/robotics.hpp:83-106.  screws: njoints records of 6 (omega; v);
 * home: one SE3 (16); q: njoints per instance; T (16) / Js (6*njoints) per
 * instance, either may be NULL. */
int ppxo_poe_fk_jacobian(const double *screws, const double *home, int njoints,
                         const double *q, double *T, double *Js,
                         size_t count, int nthreads);

/* linalg.hpp:535-603, 1189-1200.  status uses StatusCode's enumerator order
 * (linalg.hpp:362-368): 0 NORMAL 1 CONVERGED 2 DIVERGED 3 SINGULAR. */
int ppxo_linsolve_qr(int m, int n, const double *A, const double *b, double *x,
                     signed char *status, size_t count, int nthreads);
/* linalg.hpp:448-533, 1176-1187 */
int ppxo_linsolve_lu(int n, const double *A, const double *b, double *x,
                     signed char *status, size_t count, int nthreads);
/* MatrixS::I() / det(), linalg.hpp:933-965 */
int ppxo_inverse(int n, const double *A, double *Ai, size_t count, int nthreads);
int ppxo_det(int n, const double *A, double *det, size_t count, int nthreads);
/* linalg.hpp:605-924: u (m x n), w (n, unsorted), v (n x n), A = u diag(w) v^T */
int ppxo_svd(int m, int n, const double *A, double *U, double *w, double *V,
             size_t count, int nthreads);
/* linalg.hpp:886-913, 1202-1209 */
int ppxo_linsolve_svd(int m, int n, const double *A, const double *b, double *x,
                      size_t count, int nthreads);
/* pinv (n x m out), matrix norm2 = sigma_max: linalg.hpp:1211-1227, 926-931 */
int ppxo_pinv(int m, int n, const double *A, double *P, size_t count, int nthreads);
int ppxo_norm2(int m, int n, const double *A, double *s, size_t count, int nthreads);
/* linalg.hpp:967-1167: d (n), z (n x n eigenvectors in columns, may be NULL) */
int ppxo_eig_sym(int n, const double *S, double *d, double *z, int sorted,
                 size_t count, int nthreads);

/* One Newton step of test/lie_test.cpp:104-110 on demo/robotics.hpp's
 * kinematics: Vs = Ad(T(q)) log(T(q)^-1 Ttarget); q_out = q + pinv(Js(q)) Vs.
 * Vs (6 per instance) may be NULL. */
int ppxo_ik_newton_step(const double *screws, const double *home, int njoints,
                        const double *q, const double *Ttarget, double *q_out,
                        double *Vs, size_t count, int nthreads);
/* The whole loop of test/lie_test.cpp:96-113 (<= max_iter steps, stop when
 * |Vs.w| and |Vs.v| < EPS_SP *before* the last update, as the reference does).
 * iters (int per instance) may be NULL. */
int ppxo_ik_solve(const double *screws, const double *home, int njoints,
                  const double *q0, const double *Ttarget, int max_iter,
                  double *q_out, int *iters, size_t count, int nthreads);

/* kinematics<N>::inverseSpace(pose, init) of demo/robotics.hpp:126-152: bounded trust-region dogleg (details::CoDo<N,6>,
 * include/optimization.hpp:497-829) on f(q) = log(T(q) pose^-1) with the analytic Jacobian of :134-140 and the joint
 * ranges lo/hi (njoints each).  q_out = result.x if CONVERGED else init (as inverseSpace returns it); x_raw (the
 * optimiser's last iterate), fnorm (OptResult::y = 0.5 |f|) and status (CONVERGED / DIVERGED) may be NULL. */
int ppxo_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                 const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                 signed char *status, size_t count, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
