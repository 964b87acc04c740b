/*
 * ppx_cuda.h -- C ABI of libppx_cuda.so, the B200 (sm_100a) batched engine for ppx's
 * one data-parallel hot path: millions of independent fixed-size small-matrix problems
 * (so3/se3 exp & log, SE3 products/adjoints, product-of-exponentials kinematics,
 * linsolve<QR|SVD>, svdcmp, symmetric eig, IK Newton step) evaluated in one launch.
 *
 * The reference (Xtinc/matrix, "ppx") has no plugin/FFI boundary: the path is reached by
 * inline template calls on values.  Each entry point below therefore names the reference
 * routine it replaces (file:line under the reference tree) and evaluates it for `n`
 * independent instances.
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  All data pointers are DEVICE pointers owned by the
 *    caller (the *_host variants at the end take HOST pointers and stage the copies).
 *  - `layout` selects how a batch of records is laid out:
 *      PPX_LAYOUT_AOS  dense reference records back to back, i.e. exactly
 *                      std::vector<ppx::MatrixS<M,N>>::data(): instance i, element (r,c) at
 *                      base[i*M*N + r + c*M]  (column-major, include/matrixs.hpp:572-582)
 *      PPX_LAYOUT_SOA  component-major: base[(r + c*M)*ld + i], ld >= n is the component
 *                      stride in doubles (`ld` is ignored for AOS).  This is the fast layout.
 *    Every array of one call uses the same layout and the same ld.
 *  - Work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = default stream) of
 *    the CURRENT device; calls are asynchronous, never synchronise and may be issued concurrently
 *    from several host threads / devices.  The compute entry points keep no state between calls.
 *    The only process-wide settings are the two configuration knobs declared below
 *    (ppx_cuda_set_square_solve_lu, a measurement switch, and ppx_cuda_host_set_devices, the device
 *    list of the host-buffer face) and the library's caches (occupancy, streams, scratch pool).
 *  - Return value: PPX_OK (0), a PPX_ERR_* code (< 0), or a positive cudaError_t from the
 *    launch.  ppx_cuda_error_string() explains either.
 *  - Per-instance numerical status uses ppx::StatusCode's enumerator order
 *    (include/linalg.hpp:362-368): 0 NORMAL, 1 CONVERGED, 2 DIVERGED, 3 SINGULAR.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PPX_CUDA_H
#define PPX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPX_LAYOUT_AOS 0
#define PPX_LAYOUT_SOA 1

#define PPX_OK 0
#define PPX_ERR_BAD_LAYOUT (-1)
#define PPX_ERR_BAD_SHAPE (-2)      /* (m, n) / njoints not compiled into this library */
#define PPX_ERR_BAD_ARGUMENT (-3)   /* NULL pointer, ld < n, misaligned base, ... */
#define PPX_ERR_NO_DEVICE (-4)
#define PPX_ERR_HOST (-5)           /* host-side resource failure inside the library (worker thread, host memory) */

#define PPX_STATUS_NORMAL 0
#define PPX_STATUS_CONVERGED 1
#define PPX_STATUS_DIVERGED 2
#define PPX_STATUS_SINGULAR 3

#define PPX_MAX_JOINTS 8

/* ---- library / device helpers ------------------------------------------------------------- */
const char *ppx_cuda_version(void);
const char *ppx_cuda_error_string(int code);
int ppx_cuda_device_count(void);
int ppx_cuda_set_device(int device);
int ppx_cuda_malloc(void **dptr, size_t bytes);
int ppx_cuda_free(void *dptr);
/* pinned host memory, usable by every device of the process.  Blocks >= 64 MiB are mapped with transparent huge pages
 * (PPX_HOST_HUGEPAGES=0: off), faulted in by several threads and page-locked with one cudaHostRegister; smaller ones
 * come from cudaHostAlloc (PPX_HOST_ALLOC=cuda: always). */
int ppx_cuda_malloc_host(void **hptr, size_t bytes);
int ppx_cuda_free_host(void *hptr);
int ppx_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int ppx_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int ppx_cuda_stream_synchronize(void *stream);
/* number of kernel launches this library has issued in this process (all threads) */
uint64_t ppx_cuda_launch_count(void);
/* Square linsolve<SVD> systems (ppx_cuda_batch_linsolve_svd with m == n, the Newton step of ik_newton_step / ik_solve,
 * square pinv) are solved by LU with partial pivoting whenever a RIGOROUS upper bound of cond_inf(A), computed from the
 * factors (|A|_inf |M(U)^-1 e|_inf |M(L)^-1 e|_inf, M = comparison matrix), is below 1e10 -- the truncation of SVD::solve
 * (linalg.hpp:889-901) only acts beyond cond ~ 2e15, so A^+ b = A^-1 b there -- and by the Jacobi SVD otherwise.
 * enable = 0 forces the Jacobi SVD solve everywhere (a measurement switch); returns the previous setting.
 * Process-wide, read at launch time. */
int ppx_cuda_set_square_solve_lu(int enable);

/* records of `width` doubles: dense AoS (n*width) <-> component-major SoA (width*ld) */
int ppx_cuda_aos_to_soa(const double *aos, double *soa, int width, size_t n, size_t ld, void *stream);
int ppx_cuda_soa_to_aos(const double *soa, double *aos, int width, size_t n, size_t ld, void *stream);
/* out[j] = lo + (hi-lo) * u01(seed, first + j): the counter-based generator of matrix_b200/synth.py */
int ppx_cuda_fill_uniform(double *out, size_t count, uint64_t seed, uint64_t first, double lo, double hi,
                          void *stream);
/* runs a dependent-free DFMA loop on every SM and returns achieved fp64 TFLOP/s (2 flop per DFMA)
 * in *tflops; synchronises the stream (measurement helper, not part of the hot path). */
int ppx_cuda_fp64_peak_probe(double *tflops, void *stream);

/* measurement helper: writes `planes` SoA component planes of n doubles (mode 0: 64-bit stores, one instance per
 * thread, the pattern of the batch kernels; mode 1: 128-bit stores, two instances per thread) */
int ppx_cuda_probe_write(double *out, int planes, size_t n, size_t ld, int mode, void *stream);

/* ---- MatrixS<M,N>::operator*, include/matrixs.hpp:596-640 ------------------------------------ */
/* C (m x l) = A (m x n) B (n x l), zero-initialised, accumulated in the reference's k-j-i order with separately
 * rounded products and sums (no FMA): bit-identical to the reference compiled without contraction.  Unrolled kernels
 * for (3,3,3) (4,4,4) (6,6,6) (6,6,1) (4,4,1) (3,3,1); every other shape with m, n, l <= 8 runs a generic kernel. */
int ppx_cuda_batch_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count, int layout,
                          size_t ld, void *stream);

/* ---- Lie group / algebra ------------------------------------------------------------------- */
/* so3::exp, include/liegroup.hpp:116-133.  w: 3 -> R: 9 */
int ppx_cuda_batch_so3_exp(const double *w, double *R, size_t n, int layout, size_t ld, void *stream);
/* SO3::log, include/liegroup.hpp:60-92.  R: 9 -> w: 3 */
int ppx_cuda_batch_SO3_log(const double *R, double *w, size_t n, int layout, size_t ld, void *stream);
/* se3::exp, include/liegroup.hpp:253-269.  xi = (omega; v): 6 -> T: 16 */
int ppx_cuda_batch_se3_exp(const double *xi, double *T, size_t n, int layout, size_t ld, void *stream);
/* SE3::log, include/liegroup.hpp:231-250.  T: 16 -> xi: 6 */
int ppx_cuda_batch_SE3_log(const double *T, double *xi, size_t n, int layout, size_t ld, void *stream);
/* fused se3::exp followed by SE3::log (BASELINE config 1), 6 -> 6, no SE3 round trip through memory */
int ppx_cuda_batch_se3_exp_log(const double *xi, double *xi_out, size_t n, int layout, size_t ld, void *stream);
/* SE3::operator*, include/liegroup.hpp:204-207.  A, B: 16 -> C: 16 */
int ppx_cuda_batch_SE3_mul(const double *A, const double *B, double *C, size_t n, int layout, size_t ld,
                           void *stream);
/* SE3::I, include/liegroup.hpp:227-230.  T: 16 -> 16 */
int ppx_cuda_batch_SE3_inv(const double *T, double *Ti, size_t n, int layout, size_t ld, void *stream);
/* SE3::Adt, include/liegroup.hpp:217-226.  T: 16 -> Ad: 36 (6x6) */
int ppx_cuda_batch_SE3_Adt(const double *T, double *Ad, size_t n, int layout, size_t ld, void *stream);
/* se3::adt, include/liegroup.hpp:271-280.  xi: 6 -> ad: 36 (6x6) */
int ppx_cuda_batch_se3_adt(const double *xi, double *ad, size_t n, int layout, size_t ld, void *stream);
/* so3 Jacobians, include/liegroup.hpp:142-184.  which: 0 ljac, 1 ljacinv, 2 rjac, 3 rjacinv.  w: 3 -> J: 9 */
int ppx_cuda_batch_so3_jac(int which, const double *w, double *J, size_t n, int layout, size_t ld, void *stream);

/* ---- product-of-exponentials kinematics (demo/robotics.hpp) -------------------------------- */
/* kinematics<N>::forwardSpace(q) (:83-92) and ::jacobiSpace(q) (:94-106) in one pass.
 * screws: HOST pointer, njoints records (omega; v) -- kinematics<N>::JointList()[i].screw;
 * home:   HOST pointer, one SE3 (16, column-major) -- JointList().back().pose;
 * q: njoints per instance;  T: 16 per instance or NULL;  Js: 6*njoints per instance or NULL. */
int ppx_cuda_batch_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q,
                                   double *T, double *Js, size_t n, int layout, size_t ld, void *stream);
/* The three IK entry points below have kernels of their own for 6-joint chains whose joint axes follow one of three
 * patterns of unit coordinate axes (signs free): z y y y z y (the Universal-Robots family, the reference's fixture),
 * z y y x y x and z y y z y z (a spherical wrist, forearm level or upright in the home pose) -- the per-joint branch of the generic kernels is then resolved at
 * compile time (Newton step +17 %, trust region +22 %, results bitwise equal for the Newton step and loop).  Every
 * other chain, and every chain when the environment variable PPX_IK_PATTERNS is 0, runs the generic kernels. */
/* One Newton step of the reference's IK loop (test/lie_test.cpp:104-110):
 *   Vs = Ad(T(q)) log(T(q)^-1 Ttarget);  q_out = q + linsolve<SVD>(Js(q), Vs).x
 * q, q_out: njoints (<= 6);  Ttarget: 16;  Vs: 6 per instance or NULL. */
int ppx_cuda_batch_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                                  const double *Ttarget, double *q_out, double *Vs, size_t n, int layout,
                                  size_t ld, void *stream);
/* The whole loop of test/lie_test.cpp:96-113 on the device: at most max_iter steps, an instance stops
 * after the step in which |Vs.w| and |Vs.v| were both < EPS_SP.  iters (int32 per instance, dense) and
 * status (int8: CONVERGED / DIVERGED) may be NULL. */
int ppx_cuda_batch_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                            const double *Ttarget, int max_iter, double *q_out, int32_t *iters, int8_t *status,
                            size_t n, int layout, size_t ld, void *stream);

/* kinematics<6>::inverseSpace(pose, init) of demo/robotics.hpp:126-152: the bounded trust-region dogleg
 * details::CoDo<6,6> (include/optimization.hpp:497-829) on f(q) = log(T(q) pose^-1) with the joints' ranges lo/hi
 * (HOST pointers, njoints each) as the box.  q_out = the optimiser's result if CONVERGED, else init (what inverseSpace
 * returns).  Optional outputs: x_raw (last iterate, njoints), fnorm (OptResult::y = 0.5 |f|, dense double[n]),
 * status (int8 dense: CONVERGED / DIVERGED). */
int ppx_cuda_batch_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                           const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                           int8_t *status, size_t n, int layout, size_t ld, void *stream);

/* ---- dense factorisations (include/linalg.hpp) --------------------------------------------- */
/* linsolve<Factorization::QR>(A, b), :535-603, 1189-1200.  A: m*n, b: m -> x: n, status: int8 dense [n_inst]
 * (may be NULL).  On SINGULAR x = first n entries of b, as in the reference. */
int ppx_cuda_batch_linsolve_qr(int m, int n, const double *A, const double *b, double *x, int8_t *status,
                               size_t count, int layout, size_t ld, void *stream);
/* SVD<m,n> (the documented svdcmp), :605-884.  A: m*n (m >= n) -> U: m*n, w: n (>= 0, NOT sorted, like the
 * reference's), V: n*n with A = U diag(w) V^T.  U and/or V may be NULL.  One-sided Jacobi: U, w, V agree with
 * the reference's up to the order/sign freedom of the decomposition (compare sorted w and U diag(w) V^T). */
int ppx_cuda_batch_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count, int layout,
                       size_t ld, void *stream);
/* linsolve<Factorization::SVD>(A, b), :886-913, 1202-1209 (truncation 0.5*sqrt(m+n+1)*w_max*EPS_DP) */
int ppx_cuda_batch_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count,
                                int layout, size_t ld, void *stream);
/* EigenValue<n>(S, sorted), :967-1167 (the documented eig<eigensystem::Sym*>).  S: n*n (lower triangle
 * read) -> d: n (ascending when sorted), z: n*n eigenvectors in columns or NULL (SymOnlyVal). */
int ppx_cuda_batch_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count, int layout,
                           size_t ld, void *stream);

/* linsolve<Factorization::LU>(A, b), :448-533, 1176-1187 (partial pivoting; SINGULAR when a pivot < EPS_DP, x = b) */
int ppx_cuda_batch_linsolve_lu(int n, const double *A, const double *b, double *x, int8_t *status, size_t count,
                               int layout, size_t ld, void *stream);
/* MatrixS::I() and MatrixS::det() through LU, :933-965.  Ai: n*n (all zero when singular); det: 1 double */
int ppx_cuda_batch_inverse(int n, const double *A, double *Ai, size_t count, int layout, size_t ld, void *stream);
int ppx_cuda_batch_det(int n, const double *A, double *det, size_t count, int layout, size_t ld, void *stream);
/* pinv(A), :1211-1227: A m*n -> P n*m, same truncation as SVD::solve */
int ppx_cuda_batch_pinv(int m, int n, const double *A, double *P, size_t count, int layout, size_t ld, void *stream);
/* matrix norm2(A) = largest singular value, :926-931: A m*n (m < n allowed) -> 1 double */
int ppx_cuda_batch_norm2(int m, int n, const double *A, double *s, size_t count, int layout, size_t ld, void *stream);

/* ---- host-buffer face: the same operations on HOST arrays of dense reference records (AoS), i.e. on
 * std::vector<ppx::MatrixS<M,N>>::data().  Each call cuts the batch into chunks that rotate over three
 * streams so the H2D copy, the kernel and the D2H copy of neighbouring chunks overlap.  Blocking.  Pinned host
 * memory (ppx_cuda_malloc_host) lets the copies run asynchronously; pageable memory works, slower.
 * The streams and up to PPX_HOST_POOL_MIB (environment, default 1024) MiB of device scratch per device are kept
 * between calls, which is what makes small calls cheap; ppx_cuda_host_trim() returns the scratch to the driver.
 * Thread-safe.  Optional outputs (T/Js, Vs, iters, status, U, V, z) may be NULL exactly as in the device entry points. */
int ppx_cuda_host_trim(void);
/* Devices the ppx_cuda_host_* calls spread a batch over (the path shards by instance index, no exchange
 * step).  count > 0: exactly these device ordinals; count < 0: every visible device; count == 0: the calling thread's
 * current device only, which is also the default unless the environment variable PPX_CUDA_DEVICES ("all" or
 * "0,1,2,...") says otherwise.  With several devices a call runs one worker thread per device and hands chunks of the
 * batch out from one shared queue, so the result is bitwise the one-device result whatever the split.  Process-wide. */
int ppx_cuda_host_set_devices(const int *devices, int count);
/* -> number of configured devices (0 = current device only; < 0: a PPX_ERR_* code); the first `capacity` ordinals are
 * written to devices */
int ppx_cuda_host_get_devices(int *devices, int capacity);
/* measurement helper: all configured devices copy at once through `host` (bytes, pinned) for `seconds`;
 * direction 0 H2D, 1 D2H, 2 both.  gbs[i] = GB/s device i sustained meanwhile: the ceiling of any host-buffer path. */
int ppx_cuda_host_copy_probe(void *host, size_t bytes, int direction, double seconds, double *gbs);
/* Chooses, among `candidates` (count ordinals; NULL / 0: every visible device), the devices whose host links carry the
 * most when used together -- on multi-socket hosts the links interfere, and fewer devices can move more -- by greedy
 * selection on measured copy rates (`seconds` per probe, 0.1-0.2 is plenty; direction as in the copy probe: choose the
 * one that dominates the calls to come), and makes them the device list.  `host` (bytes, pinned) is scratch for the
 * copies.  *aggregate_gbs (may be NULL) = what the chosen set sustains.  Read the choice with ppx_cuda_host_get_devices. */
int ppx_cuda_host_tune_devices(const int *candidates, int count, void *host, size_t bytes, int direction, double seconds,
                               double *aggregate_gbs);
int ppx_cuda_host_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count);
int ppx_cuda_host_so3_exp(const double *w, double *R, size_t n);
int ppx_cuda_host_SO3_log(const double *R, double *w, size_t n);
int ppx_cuda_host_se3_exp(const double *xi, double *T, size_t n);
int ppx_cuda_host_SE3_log(const double *T, double *xi, size_t n);
int ppx_cuda_host_se3_exp_log(const double *xi, double *xi_out, size_t n);
int ppx_cuda_host_SE3_mul(const double *A, const double *B, double *C, size_t n);
int ppx_cuda_host_SE3_inv(const double *T, double *Ti, size_t n);
int ppx_cuda_host_SE3_Adt(const double *T, double *Ad, size_t n);
int ppx_cuda_host_se3_adt(const double *xi, double *ad, size_t n);
int ppx_cuda_host_so3_jac(int which, const double *w, double *J, size_t n);
int ppx_cuda_host_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q,
                                  double *T, double *Js, size_t n);
int ppx_cuda_host_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                                 const double *Ttarget, double *q_out, double *Vs, size_t n);
int ppx_cuda_host_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                           const double *Ttarget, int max_iter, double *q_out, int32_t *iters, int8_t *status,
                           size_t n);
int ppx_cuda_host_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                          const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                          int8_t *status, size_t n);
int ppx_cuda_host_linsolve_qr(int m, int n, const double *A, const double *b, double *x, int8_t *status,
                              size_t count);
int ppx_cuda_host_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count);
int ppx_cuda_host_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count);
int ppx_cuda_host_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count);
int ppx_cuda_host_linsolve_lu(int n, const double *A, const double *b, double *x, int8_t *status, size_t count);
int ppx_cuda_host_inverse(int n, const double *A, double *Ai, size_t count);
int ppx_cuda_host_det(int n, const double *A, double *det, size_t count);
int ppx_cuda_host_pinv(int m, int n, const double *A, double *P, size_t count);
int ppx_cuda_host_norm2(int m, int n, const double *A, double *s, size_t count);

#ifdef __cplusplus
}
#endif
#endif
