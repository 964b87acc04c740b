// host.cu -- the host-buffer face of the C ABI (ppx_cuda_host_*): what a ppx user holding
// std::vector<ppx::MatrixS<M,N>> calls.  Inputs and outputs are HOST arrays of dense reference
// records (AoS); the batch is cut into chunks that rotate over a few streams so that the H2D copy of
// chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full duplex and,
// at ~55 GB/s against 6.5 TB/s of HBM, the only thing that matters end to end).  Device scratch comes
// from the stream-ordered allocator and is released before the call returns.  Pinned host memory
// (ppx_cuda_malloc_host) makes the copies truly asynchronous; pageable memory works but serialises.
// Streams and device scratch are kept between calls (per device: a few non-blocking streams and a memory pool that
// retains up to PPX_HOST_POOL_MIB, default 1024, of freed scratch), so a small call costs tens of microseconds instead
// of the ~0.7 ms that creating a stream and mapping fresh device memory took; ppx_cuda_host_trim() gives it back.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "ppx_launch.cuh"

namespace
{

struct HostIn
{
    const void *ptr;
    size_t bytes; // per instance
};
struct HostOut
{
    void *ptr; // may be NULL (optional output)
    size_t bytes;
};

constexpr int MAX_SLOTS = 8;
// defaults: 3 chunks in flight, 48 MiB of traffic on the widest side per chunk; PPX_HOST_SLOTS / PPX_HOST_CHUNK_MIB
// override them (tuning knobs, read once)
inline int env_slots()
{
    static const int v = [] {
        const char *e = std::getenv("PPX_HOST_SLOTS");
        const int n = e ? std::atoi(e) : 3;
        return n < 1 ? 1 : (n > MAX_SLOTS ? MAX_SLOTS : n);
    }();
    return v;
}
inline size_t env_chunk_bytes()
{
    static const size_t v = [] {
        const char *e = std::getenv("PPX_HOST_CHUNK_MIB");
        const long n = e ? std::atol(e) : 48;
        return (size_t)(n < 1 ? 1 : n) << 20;
    }();
    return v;
}

// per-device resources kept between calls
struct DevCtx
{
    std::vector<cudaStream_t> idle; // streams not in use by a call
};
std::mutex g_host_mu;
DevCtx g_host_ctx[64];

// -> `want` streams and the scratch pool of the current device
int host_acquire(int want, cudaStream_t *st, cudaMemPool_t &pool, int &device)
{
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return (int)e;
    if (device < 0 || device >= 64)
        return PPX_ERR_NO_DEVICE;
    const int prc = ppx_scratch_pool(device, &pool);
    if (prc != PPX_OK)
        return prc;
    std::lock_guard<std::mutex> lock(g_host_mu);
    DevCtx &c = g_host_ctx[device];
    for (int s = 0; s < want; s++)
    {
        if (!c.idle.empty())
        {
            st[s] = c.idle.back();
            c.idle.pop_back();
        }
        else
        {
            e = cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
            if (e != cudaSuccess)
            {
                for (int k = 0; k < s; k++)
                    c.idle.push_back(st[k]);
                return (int)e;
            }
        }
    }
    return PPX_OK;
}
void host_release(int device, int n, cudaStream_t *st)
{
    std::lock_guard<std::mutex> lock(g_host_mu);
    for (int s = 0; s < n; s++)
        if (st[s])
            g_host_ctx[device].idle.push_back(st[s]);
}

#define PPX_TRY(expr)                 \
    do                                \
    {                                 \
        cudaError_t e_ = (expr);      \
        if (e_ != cudaSuccess)        \
        {                             \
            rc = (int)e_;             \
            goto done;                \
        }                             \
    } while (0)

// run(count, d_in[], d_out[], stream) launches the AoS kernel for one chunk
template <class Run>
int pipeline(const std::vector<HostIn> &ins, const std::vector<HostOut> &outs, size_t n, Run run)
{
    if (n == 0)
        return PPX_OK;
    for (const HostIn &a : ins)
        if (a.ptr == nullptr)
            return PPX_ERR_BAD_ARGUMENT;
    size_t in_b = 0, out_b = 0;
    for (const HostIn &a : ins)
        in_b += a.bytes;
    for (const HostOut &a : outs)
        out_b += a.ptr ? a.bytes : 0;
    const size_t widest = std::max<size_t>(std::max(in_b, out_b), 1);
    size_t chunk = std::max<size_t>(env_chunk_bytes() / widest, 1024);
    chunk = (chunk + 1023) / 1024 * 1024;
    chunk = std::min(chunk, n);

    int rc = PPX_OK;
    const int SLOTS = env_slots();
    cudaStream_t st[MAX_SLOTS] = {};
    std::vector<void *> d_in(SLOTS * ins.size(), nullptr), d_out(SLOTS * outs.size(), nullptr);
    const int nslots = (int)std::min<size_t>(SLOTS, (n + chunk - 1) / chunk);
    cudaMemPool_t pool = nullptr;
    int device = 0;
    rc = host_acquire(nslots, st, pool, device);
    if (rc != PPX_OK)
        return rc;
    for (int s = 0; s < nslots; s++)
    {
        for (size_t a = 0; a < ins.size(); a++)
            PPX_TRY(cudaMallocFromPoolAsync(&d_in[s * ins.size() + a], chunk * ins[a].bytes, pool, st[s]));
        for (size_t a = 0; a < outs.size(); a++)
            if (outs[a].ptr)
                PPX_TRY(cudaMallocFromPoolAsync(&d_out[s * outs.size() + a], chunk * outs[a].bytes, pool, st[s]));
    }
    for (size_t first = 0, k = 0; first < n; first += chunk, k++)
    {
        const int s = (int)(k % nslots);
        const size_t cnt = std::min(chunk, n - first);
        for (size_t a = 0; a < ins.size(); a++)
            PPX_TRY(cudaMemcpyAsync(d_in[s * ins.size() + a], (const char *)ins[a].ptr + first * ins[a].bytes,
                                    cnt * ins[a].bytes, cudaMemcpyHostToDevice, st[s]));
        rc = run(cnt, &d_in[s * ins.size()], &d_out[s * outs.size()], (void *)st[s]);
        if (rc != PPX_OK)
            goto done;
        for (size_t a = 0; a < outs.size(); a++)
            if (outs[a].ptr)
                PPX_TRY(cudaMemcpyAsync((char *)outs[a].ptr + first * outs[a].bytes, d_out[s * outs.size() + a],
                                        cnt * outs[a].bytes, cudaMemcpyDeviceToHost, st[s]));
    }
done:
    for (int s = 0; s < nslots; s++)
    {
        for (size_t a = 0; a < ins.size(); a++)
            if (d_in[s * ins.size() + a])
                cudaFreeAsync(d_in[s * ins.size() + a], st[s]);
        for (size_t a = 0; a < outs.size(); a++)
            if (d_out[s * outs.size() + a])
                cudaFreeAsync(d_out[s * outs.size() + a], st[s]);
        cudaError_t e = cudaStreamSynchronize(st[s]);
        if (rc == PPX_OK && e != cudaSuccess)
            rc = (int)e;
    }
    host_release(device, nslots, st);
    return rc;
}

constexpr size_t D = sizeof(double);
typedef const double *cd;
typedef double *md;

} // namespace

extern "C"
{

int ppx_cuda_host_trim(void) { return ppx_scratch_trim(); }

int ppx_cuda_host_so3_exp(const double *w, double *R, size_t n)
{
    return pipeline({{w, 3 * D}}, {{R, 9 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_so3_exp((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SO3_log(const double *R, double *w, size_t n)
{
    return pipeline({{R, 9 * D}}, {{w, 3 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SO3_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_exp(const double *xi, double *T, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{T, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_exp((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_log(const double *T, double *xi, size_t n)
{
    return pipeline({{T, 16 * D}}, {{xi, 6 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_exp_log(const double *xi, double *xi_out, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{xi_out, 6 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_exp_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_mul(const double *A, const double *B, double *C, size_t n)
{
    return pipeline({{A, 16 * D}, {B, 16 * D}}, {{C, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_mul((cd)i[0], (cd)i[1], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_inv(const double *T, double *Ti, size_t n)
{
    return pipeline({{T, 16 * D}}, {{Ti, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_inv((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_Adt(const double *T, double *Ad, size_t n)
{
    return pipeline({{T, 16 * D}}, {{Ad, 36 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_Adt((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_adt(const double *xi, double *ad, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{ad, 36 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_adt((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_so3_jac(int which, const double *w, double *J, size_t n)
{
    return pipeline({{w, 3 * D}}, {{J, 9 * D}}, n, [which](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_so3_jac(which, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q,
                                  double *T, double *Js, size_t n)
{
    if (njoints < 1 || njoints > PPX_MAX_JOINTS)
        return PPX_ERR_BAD_SHAPE;
    const size_t nj = (size_t)njoints;
    return pipeline({{q, nj * D}}, {{T, 16 * D}, {Js, 6 * nj * D}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_poe_fk_jacobian(screws, home, njoints, (cd)i[0], (md)o[0], (md)o[1], c,
                                                              PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                                 const double *Ttarget, double *q_out, double *Vs, size_t n)
{
    if (njoints < 3 || njoints > 6)
        return PPX_ERR_BAD_SHAPE;
    const size_t nj = (size_t)njoints;
    return pipeline({{q, nj * D}, {Ttarget, 16 * D}}, {{q_out, nj * D}, {Vs, 6 * D}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_newton_step(screws, home, njoints, (cd)i[0], (cd)i[1], (md)o[0],
                                                             (md)o[1], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                           const double *Ttarget, int max_iter, double *q_out, int32_t *iters, int8_t *status,
                           size_t n)
{
    if (njoints != 6)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{q0, 6 * D}, {Ttarget, 16 * D}}, {{q_out, 6 * D}, {iters, sizeof(int32_t)}, {status, 1}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_solve(screws, home, njoints, (cd)i[0], (cd)i[1], max_iter, (md)o[0],
                                                       (int32_t *)o[1], (int8_t *)o[2], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_linsolve_qr(int m, int n, const double *A, const double *b, double *x, int8_t *status,
                              size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}, {b, (size_t)m * D}}, {{x, (size_t)n * D}, {status, 1}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_qr(m, n, (cd)i[0], (cd)i[1], (md)o[0], (int8_t *)o[1], c,
                                                          PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{U, (size_t)m * n * D}, {w, (size_t)n * D}, {V, (size_t)n * n * D}},
                    count, [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_svd(m, n, (cd)i[0], (md)o[0], (md)o[1], (md)o[2], c, PPX_LAYOUT_AOS, 0,
                                                  s);
                    });
}

int ppx_cuda_host_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}, {b, (size_t)m * D}}, {{x, (size_t)n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_svd(m, n, (cd)i[0], (cd)i[1], (md)o[0], c, PPX_LAYOUT_AOS, 0,
                                                           s);
                    });
}

int ppx_cuda_host_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{S, (size_t)n * n * D}}, {{d, (size_t)n * D}, {z, (size_t)n * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_eig_sym(n, (cd)i[0], (md)o[0], (md)o[1], sorted, c, PPX_LAYOUT_AOS, 0,
                                                      s);
                    });
}

int ppx_cuda_host_linsolve_lu(int n, const double *A, const double *b, double *x, int8_t *status, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}, {b, (size_t)n * D}}, {{x, (size_t)n * D}, {status, 1}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_lu(n, (cd)i[0], (cd)i[1], (md)o[0], (int8_t *)o[1], c,
                                                          PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_inverse(int n, const double *A, double *Ai, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}}, {{Ai, (size_t)n * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_inverse(n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_det(int n, const double *A, double *det, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}}, {{det, D}}, count, [=](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_det(n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_pinv(int m, int n, const double *A, double *P, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{P, (size_t)m * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_pinv(m, n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_norm2(int m, int n, const double *A, double *sv, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{sv, D}}, count, [=](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_norm2(m, n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                          const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                          int8_t *status, size_t n)
{
    if (njoints != 6)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{Ttarget, 16 * D}, {init, 6 * D}}, {{q_out, 6 * D}, {x_raw, 6 * D}, {fnorm, D}, {status, 1}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_codo(screws, home, njoints, lo, hi, (cd)i[0], (cd)i[1], (md)o[0], (md)o[1],
                                                      (md)o[2], (int8_t *)o[3], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

} // extern "C"
