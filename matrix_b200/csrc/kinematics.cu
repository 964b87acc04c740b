// kinematics.cu -- product-of-exponentials forward kinematics + space Jacobian and the IK Newton
// step built from them (demo/robotics.hpp:83-106 and test/lie_test.cpp:96-113 of the reference).
//
// The chain (screw axes + home pose) is the same for every instance of a call, so the host folds
// everything that depends on it alone into a ChainConst that travels as a __grid_constant__ kernel
// parameter (constant bank: free FMA operands, no registers).  Per instance the kernel then needs
// only NJ sincos, NJ-1 rigid products shared between the pose and the Jacobian columns, and NJ-1
// structured adjoint applications: no division, no square root, no 4x4 matrix powers.
#include <cmath>
#include <cstdlib>

#include "ppx_launch.cuh"

// ---- configuration of the inverseSpace (CoDo) kernel, needed before ppx_codo.cuh is read ----
// One CTA of 8 warps per SM (232 registers and 816 B of shared memory per lane).  PPX_CODO_ALIGN = 1 makes the warps of
// the CTA start every trip together (the loop's exit test becomes a __syncthreads_or): that paid (+4 %) while the body
// of a trip was far larger than the SM's instruction cache and the kernel waited for instruction fetches; with the
// present footprint the barrier costs more than it saves (-2 %), so it is off.
#ifndef PPX_CODO_BLOCK
#define PPX_CODO_BLOCK 256
#endif
#ifndef PPX_CODO_MINB
#define PPX_CODO_MINB 1
#endif
#ifndef PPX_CODO_ALIGN
#define PPX_CODO_ALIGN 0
#endif


#define PPX_CODO_JSTRIDE PPX_CODO_BLOCK // a lane's Jacobians are interleaved by thread in shared memory
#include "ppx_codo.cuh"
constexpr size_t CODO_SMEM = ppx_dev::CODO_LANE_DOUBLES * sizeof(double) * PPX_CODO_BLOCK;

using namespace ppx_dev;

namespace
{

// BASELINE config 2.  48 B in, 128 + 288 B out per instance at NJ = 6: write-dominated and HBM-bound.
template <int NJ, bool WANT_T, bool WANT_J>
struct OpFkJacobian
{
    // 3 CTAs of 128 per SM (168 registers): the record (AoS) layout then needs no spills around its staging tile and
    // reaches 0.94 of the HBM peak (0.86 at 4 CTAs / 128 registers); the component-plane layout is indifferent (0.905)
#ifndef PPX_FK_MINB
#define PPX_FK_MINB 3
#endif
#ifndef PPX_FK_BLOCK
#define PPX_FK_BLOCK 128
#endif
    static constexpr int BLOCK = PPX_FK_BLOCK, MINB = PPX_FK_MINB, MAXW = (WANT_J ? 6 * NJ : 16) > 16 ? (WANT_J ? 6 * NJ : 16) : 16;
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

// dq = linsolve<SVD>(Js, Vs).x: a square Jacobian (6 joints) needs no V, and no SVD either unless it is close to
// singular (square_pinv_solve).  PRECOND selects the single-precision preconditioning of the Jacobi fallback.
template <int NJ, bool PRECOND>
__device__ __forceinline__ typename std::enable_if<NJ == 6>::type svd_solve_dispatch(double (&js)[36], const double (&vs)[6],
                                                                                    double (&dq)[6], bool allow_lu)
{
    square_pinv_solve<6, PRECOND>(js, vs, dq, allow_lu);
}
template <int NJ, bool PRECOND>
__device__ __forceinline__ typename std::enable_if<NJ != 6>::type svd_solve_dispatch(double (&js)[6 * NJ],
                                                                                    const double (&vs)[6], double (&dq)[NJ],
                                                                                    bool)
{
    double V[NJ * NJ], w2[NJ];
    jacobi_svd_core<6, NJ, true>(js, V, w2);
    jacobi_svd_solve<6, NJ>(js, V, w2, vs, dq);
}

// Vs = Ad(T(q)) log(T(q)^-1 Tt),  dq = pinv(Js(q)) Vs  (test/lie_test.cpp:104-110)
template <int NJ, bool PRECOND = true, class Kinds = RuntimeKinds>
__device__ __forceinline__ void ik_newton_update(const ChainConst<NJ> &ch, const Rigid &target, double (&q)[NJ],
                                                 double (&vs)[6], bool allow_lu)
{
    Rigid t, ti, x;
    double js[6 * NJ];
    poe_fk_jacobian<NJ, true, true, Kinds>(ch, q, t, [&](int col, const double(&v)[6]) {
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
    svd_solve_dispatch<NJ, PRECOND>(js, vs, dq, allow_lu);
#pragma unroll
    for (int i = 0; i < NJ; i++)
        q[i] += dq[i];
}

// BASELINE config 5.  48 + 128 B in, 48 B out.  Two instantiations per chain length:
//   SVD_ONLY = false  the common form: for 6 joints the square solve goes through LU with partial pivoting whenever the
//                     rigorous condition bound allows (square_pinv_solve), with the Jacobi solve as the per-warp
//                     fallback; 3 CTAs per SM (168 registers), inputs of the next tile staged with cp.async.
//   SVD_ONLY = true   the Jacobi SVD solve on every instance (ppx_cuda_set_square_solve_lu(0): the step exactly as
//                     north_star words it): 2 CTAs per SM (255 registers), no single-precision preconditioning --
//                     the square solve accumulates no V, so the preconditioner's V0 costs more than it saves here.
// Measured on B200 (2^25 instances): 10.1 G inst/s / 2.13 G inst/s.  The register allocation of the common form is
// fragile -- the same arithmetic with the fallback out of line, or split into an LU pass and an SVD pass over marked
// instances, ran at 7.9-8.7 G inst/s -- so it is kept in the shape that measured best.
#ifndef PPX_IK_MINB
#define PPX_IK_MINB 3
#endif
// joint-axis patterns with kernels of their own (FixedKinds): the Universal-Robots family and its many look-alikes
// (z y y y z y: UR3/5/10/16/20, the reference's own fixture), and the classic 6R arm with a spherical wrist in its
// home pose (z y y x y x: PUMA, ABB IRB, KUKA KR, Fanuc, ...).  Any other chain runs the RuntimeKinds kernels.
using KindsUR = FixedKinds<3, 2, 2, 2, 3, 2>;
using KindsWrist = FixedKinds<3, 2, 2, 1, 2, 1>;
// ... and the same arm drawn with its forearm pointing up in the home pose (z y y z y z: Denso, Epson, Staubli style)
using KindsUpright = FixedKinds<3, 2, 2, 3, 2, 3>;

template <int NJ, bool SVD_ONLY, class Kinds = RuntimeKinds>
struct OpIkStep
{
    static constexpr int BLOCK = 128, MINB = SVD_ONLY ? 2 : PPX_IK_MINB, MAXW = 16;
    ChainConst<NJ> ch;
    const double *q;
    const double *Tt;
    double *q_out;
    double *Vs;
    // joint angles and target pose of the next tile staged (cp.async) behind the arithmetic of the current one
    static constexpr bool PREFETCH = !SVD_ONLY;
    static constexpr int IN_WIDTH = NJ + 16 + 2;
    template <class IO>
    __device__ __forceinline__ void prefetch(const IO &io, double *stage) const
    {
        io.template async_load<NJ>(q, stage);
        io.template async_load<16>(Tt, stage + IO::stage_doubles(BLOCK, NJ));
    }
    template <class IO>
    __device__ __forceinline__ void compute_store(const IO &io, const double *stage) const
    {
        double qi[NJ], m[16];
        io.template read_staged<NJ>(stage, qi);
        io.template read_staged<16>(stage + IO::stage_doubles(BLOCK, NJ), m);
        step_store(io, qi, m);
    }
    template <class IO>
    __device__ __forceinline__ void operator()(const IO &io) const
    {
        double qi[NJ], m[16];
        io.load(q, qi);
        io.load(Tt, m);
        step_store(io, qi, m);
    }
    template <class IO>
    __device__ __forceinline__ void step_store(const IO &io, double (&qi)[NJ], const double (&m)[16]) const
    {
        double vs[6];
        Rigid target;
        rigid_from_record(m, target);
        ik_newton_update<NJ, !SVD_ONLY, Kinds>(ch, target, qi, vs, !SVD_ONLY);
        io.store(q_out, qi);
        if (Vs != nullptr)
            io.store(Vs, vs);
    }
};

// dense-record / component-plane access for the persistent kernels below (one instance per lane, claimed at run time)
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

// lanes of a warp without work claim consecutive unprocessed instances from a global counter: one atomicAdd per warp
// and refill round.  -> the claimed index (>= n: nothing left); call with the whole warp converged.
__device__ __forceinline__ size_t claim_instances(bool wants, unsigned long long *next)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned want = __ballot_sync(0xffffffffu, wants);
    if (want == 0)
        return ~(size_t)0;
    const int leader = __ffs(want) - 1;
    unsigned long long first = 0;
    if ((int)lane == leader)
        first = atomicAdd(next, (unsigned long long)__popc(want));
    first = __shfl_sync(0xffffffffu, first, leader);
    return wants ? (size_t)first + __popc(want & ((1u << lane) - 1)) : ~(size_t)0;
}

// The whole loop of test/lie_test.cpp:96-113: the convergence test uses the Vs of the current step and the update of
// that step is still applied, exactly as the reference does.  Iteration counts differ from target to target (3 .. 20,
// mean 4.1 on the bench workload, but 7.3 for the slowest lane of a warp), so the kernel is persistent with per-lane
// refill: a lane whose instance has stopped stores it and takes the next one; every trip of the loop is one Newton step
// of all 32 lanes.  The step contains warp votes (Jacobi sweep termination), so lanes without work run it too, on the
// instance they finished last.
struct IkSolveArgs
{
    ChainConst<6> ch;
    const double *q0;
    const double *Tt;
    double *q_out;
    int32_t *iters;
    int8_t *status;
    int max_iter;
    bool allow_lu;
};

template <int LAYOUT, class Kinds>
__global__ void __launch_bounds__(128, PPX_IK_MINB)
    ik_solve_kernel(const __grid_constant__ IkSolveArgs a, size_t n, size_t ld, unsigned long long *next)
{
    constexpr int NJ = 6;
    double qi[NJ], vs[6];
    Rigid target;
    rigid_identity(target);
#pragma unroll
    for (int k = 0; k < NJ; k++)
        qi[k] = 0.25 * (k + 1); // a regular configuration for lanes that never get work
    size_t idx = 0;
    int it = 0;
    bool busy = false, finished = false, conv = false;
    for (;;)
    {
        const size_t got = claim_instances(!busy, next);
        if (!busy && got < n)
        {
            idx = got;
            double m[16];
#pragma unroll
            for (int k = 0; k < 16; k++)
                m[k] = rec_get<LAYOUT>(a.Tt, idx, k, 16, ld);
#pragma unroll
            for (int k = 0; k < NJ; k++)
                qi[k] = rec_get<LAYOUT>(a.q0, idx, k, NJ, ld);
            rigid_from_record(m, target);
            busy = true;
            it = 0;
            conv = false;
            finished = a.max_iter == 0;
        }
        if (!__any_sync(0xffffffffu, busy))
            break;
        double qn[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++)
            qn[i] = qi[i];
        ik_newton_update<NJ, true, Kinds>(a.ch, target, qn, vs, a.allow_lu);
        if (busy && !finished)
        {
#pragma unroll
            for (int i = 0; i < NJ; i++)
                qi[i] = qn[i];
            it++;
            // |Vs.w| < EPS_SP and |Vs.v| < EPS_SP (lie_test.cpp:103), compared as squares
            const double ew2 = fma(vs[2], vs[2], fma(vs[1], vs[1], vs[0] * vs[0]));
            const double ev2 = fma(vs[5], vs[5], fma(vs[4], vs[4], vs[3] * vs[3]));
            conv = ew2 < EPS_SP * EPS_SP && ev2 < EPS_SP * EPS_SP;
            finished = conv || it >= a.max_iter;
        }
        if (busy && finished)
        {
#pragma unroll
            for (int k = 0; k < NJ; k++)
                rec_put<LAYOUT>(a.q_out, idx, k, NJ, ld, qi[k]);
            if (a.iters != nullptr)
                a.iters[idx] = it;
            if (a.status != nullptr)
                a.status[idx] = conv ? PPX_STATUS_CONVERGED : PPX_STATUS_DIVERGED;
            busy = false;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kinematics<6>::inverseSpace(pose, init) of demo/robotics.hpp:126-152.  The optimiser itself is the
// state machine of ppx_codo.cuh; this kernel feeds it.
//
// Persistent warps with per-lane work refill: every lane owns one instance at a time and, when its optimiser returns,
// claims the next unprocessed instance from a global counter (one atomicAdd per warp per refill round).  Each trip of
// the loop is one evaluation of residual + Jacobian for all 32 lanes, whatever phase of the optimiser they are in.
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

template <int LAYOUT, class Kinds>
__global__ void __launch_bounds__(PPX_CODO_BLOCK, PPX_CODO_MINB)
    ik_codo_kernel(const __grid_constant__ IkCodoArgs a, size_t n, size_t ld, unsigned long long *next)
{
    CodoLane st;
    extern __shared__ __align__(16) double codo_smem[];
    codo_attach(st, LaneArray{codo_smem + threadIdx.x});
    Rigid poseI; // pose^-1
    double q0[6];
    size_t idx = 0;
    bool busy = false, finished = false;
    for (;;)
    {
        // refill: lanes without work claim consecutive instances
        const size_t got = claim_instances(!busy, next);
        if (!busy && got < n)
        {
            idx = got;
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
            rigid_inv(pose, poseI);
            busy = true;
            finished = codo_begin(a.lo, a.hi, st);
        }
#if PPX_CODO_ALIGN
        if (!__syncthreads_or(busy)) // also the barrier that lines the warps up
            break;
#else
        if (!__any_sync(0xffffffffu, busy))
            break;
#endif
        const unsigned active = __ballot_sync(0xffffffffu, busy && !finished);
        if (busy && !finished)
        {
            double fp[6];
            codo_eval<6, Kinds>(a.ch, poseI, st.probe, fp, codo_jac_probe(st));
            finished = codo_advance(active, a.lo, a.hi, st, fp);
        }
        if (busy && finished)
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
            finished = false;
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

// PPX_IK_PATTERNS=0 in the environment keeps every chain on the RuntimeKinds kernels (tests compare the two)
inline bool ik_patterns_enabled()
{
    const char *e = std::getenv("PPX_IK_PATTERNS");
    return !(e && e[0] == '0');
}

template <int NJ, class Kinds>
int ik_step_launch(const ChainConst<NJ> &ch, bool lu, const double *q, const double *Tt, double *q_out, double *Vs,
                   size_t n, int layout, size_t ld, void *stream)
{
    if (lu)
        return dispatch(layout, OpIkStep<NJ, false, Kinds>{ch, q, Tt, q_out, Vs}, n, ld, stream);
    return dispatch(layout, OpIkStep<NJ, true, Kinds>{ch, q, Tt, q_out, Vs}, n, ld, stream);
}
template <int NJ>
typename std::enable_if<NJ == 6, int>::type ik_step_dispatch(const ChainConst<NJ> &ch, bool lu, const double *q,
                                                             const double *Tt, double *q_out, double *Vs, size_t n,
                                                             int layout, size_t ld, void *stream)
{
    if (ik_patterns_enabled())
    {
        if (KindsUR::matches(ch))
            return ik_step_launch<NJ, KindsUR>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
        if (KindsWrist::matches(ch))
            return ik_step_launch<NJ, KindsWrist>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
        if (KindsUpright::matches(ch))
            return ik_step_launch<NJ, KindsUpright>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
    }
    return ik_step_launch<NJ, RuntimeKinds>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
}
template <int NJ>
typename std::enable_if<NJ != 6, int>::type ik_step_dispatch(const ChainConst<NJ> &ch, bool lu, const double *q,
                                                             const double *Tt, double *q_out, double *Vs, size_t n,
                                                             int layout, size_t ld, void *stream)
{
    return ik_step_launch<NJ, RuntimeKinds>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
}

template <int NJ>
int ik_step_n(const double *screws, const double *home, const double *q, const double *Tt, double *q_out,
              double *Vs, size_t n, int layout, size_t ld, void *stream)
{
    ChainConst<NJ> ch;
    fold_chain<NJ>(screws, home, ch);
    const bool lu = NJ == 6 && g_ppx_square_lu.load(std::memory_order_relaxed) != 0;
    return ik_step_dispatch<NJ>(ch, lu, q, Tt, q_out, Vs, n, layout, ld, stream);
}

// Launch of a persistent kernel with a work counter: as many CTAs as are resident at once; the counter lives for this
// launch only (stream-ordered allocation from the library's retained pool, zeroed on the stream).
template <auto Kern, class Args>
int launch_persistent(int block, size_t smem, const Args &a, size_t n, size_t ld, void *stream)
{
    DeviceInfo info;
    int device = 0;
    int rc = device_info(info, device);
    if (rc != PPX_OK)
        return rc;
    static std::atomic<int> blocks_per_sm[64]; // per instantiation (= per kernel), per device; 0 = not asked yet
    int per_sm = blocks_per_sm[device].load(std::memory_order_relaxed);
    if (per_sm == 0)
    {
        if (smem > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return (int)e;
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, Kern, block, smem);
        if (e != cudaSuccess)
            return (int)e;
        per_sm = per_sm > 0 ? per_sm : 1;
        blocks_per_sm[device].store(per_sm, std::memory_order_relaxed);
    }
    cudaMemPool_t pool = nullptr;
    rc = ppx_scratch_pool(device, &pool);
    if (rc != PPX_OK)
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *next = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync((void **)&next, sizeof(*next), pool, s);
    if (e != cudaSuccess)
        return (int)e;
    e = cudaMemsetAsync(next, 0, sizeof(*next), s);
    if (e == cudaSuccess)
    {
        const size_t want = (n + block - 1) / block;
        const size_t cap = (size_t)info.sms * per_sm;
        Kern<<<(unsigned)(want < cap ? want : cap), block, smem, s>>>(a, n, ld, next);
        g_ppx_launch_count.fetch_add(1, std::memory_order_relaxed);
        e = cudaPeekAtLastError();
    }
    const cudaError_t fe = cudaFreeAsync(next, s);
    return (int)(e != cudaSuccess ? e : fe);
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
    if (layout != PPX_LAYOUT_AOS && layout != PPX_LAYOUT_SOA)
        return PPX_ERR_BAD_LAYOUT;
    if (layout == PPX_LAYOUT_SOA && ld < n)
        return PPX_ERR_BAD_ARGUMENT;
    IkSolveArgs a;
    fold_chain<6>(screws, home, a.ch);
    a.q0 = q0, a.Tt = Ttarget, a.q_out = q_out, a.iters = iters, a.status = status, a.max_iter = max_iter;
    a.allow_lu = g_ppx_square_lu.load(std::memory_order_relaxed) != 0;
    const bool aos = layout == PPX_LAYOUT_AOS;
    if (ik_patterns_enabled() && KindsUR::matches(a.ch))
        return aos ? launch_persistent<ik_solve_kernel<LAYOUT_AOS, KindsUR>>(128, 0, a, n, ld, stream)
                   : launch_persistent<ik_solve_kernel<LAYOUT_SOA, KindsUR>>(128, 0, a, n, ld, stream);
    if (ik_patterns_enabled() && KindsWrist::matches(a.ch))
        return aos ? launch_persistent<ik_solve_kernel<LAYOUT_AOS, KindsWrist>>(128, 0, a, n, ld, stream)
                   : launch_persistent<ik_solve_kernel<LAYOUT_SOA, KindsWrist>>(128, 0, a, n, ld, stream);
    if (ik_patterns_enabled() && KindsUpright::matches(a.ch))
        return aos ? launch_persistent<ik_solve_kernel<LAYOUT_AOS, KindsUpright>>(128, 0, a, n, ld, stream)
                   : launch_persistent<ik_solve_kernel<LAYOUT_SOA, KindsUpright>>(128, 0, a, n, ld, stream);
    return aos ? launch_persistent<ik_solve_kernel<LAYOUT_AOS, RuntimeKinds>>(128, 0, a, n, ld, stream)
               : launch_persistent<ik_solve_kernel<LAYOUT_SOA, RuntimeKinds>>(128, 0, a, n, ld, stream);
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
    const bool aos = layout == PPX_LAYOUT_AOS;
    if (ik_patterns_enabled() && KindsUR::matches(a.ch))
        return aos ? launch_persistent<ik_codo_kernel<LAYOUT_AOS, KindsUR>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream)
                   : launch_persistent<ik_codo_kernel<LAYOUT_SOA, KindsUR>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream);
    if (ik_patterns_enabled() && KindsWrist::matches(a.ch))
        return aos ? launch_persistent<ik_codo_kernel<LAYOUT_AOS, KindsWrist>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream)
                   : launch_persistent<ik_codo_kernel<LAYOUT_SOA, KindsWrist>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream);
    if (ik_patterns_enabled() && KindsUpright::matches(a.ch))
        return aos ? launch_persistent<ik_codo_kernel<LAYOUT_AOS, KindsUpright>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream)
                   : launch_persistent<ik_codo_kernel<LAYOUT_SOA, KindsUpright>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream);
    return aos ? launch_persistent<ik_codo_kernel<LAYOUT_AOS, RuntimeKinds>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream)
               : launch_persistent<ik_codo_kernel<LAYOUT_SOA, RuntimeKinds>>(PPX_CODO_BLOCK, CODO_SMEM, a, n, ld, stream);
}

} // extern "C"
