// codo_host.cpp -- TEST HARNESS: the device state machine of matrix_b200/csrc/ppx_codo.cuh compiled for the host.
//
// ppx_codo.cuh (and the parts of ppx_math.cuh it uses) is plain C++ apart from the CUDA function qualifiers, so the
// very code the GPU runs per lane can be replayed here, lane by lane, against the reference's fixtures
// (tests/golden/ref_outputs.npz, codo.*) in the CPU test-suite: a logic error in the state machine shows up without a
// GPU.  This is not a product path: nothing in matrix_b200/ or include/ builds, links or loads this file; it is
// compiled by tests/test_codo_host.py into a scratch .so next to the test.
//
//   g++ -O2 -std=c++17 -shared -fPIC -ffp-contract=off -Wno-unknown-pragmas tests/cpp/codo_host.cpp -o codo_host.so
#include "cuda_host_shims.h"

#include "../../matrix_b200/csrc/ppx_codo.cuh"

using namespace ppx_dev;

// Same contract as ppx_cuda_batch_ik_codo on dense records; evals (optional) = evaluations of residual + Jacobian.
extern "C" int codo_host_run(const double *screws, const double *home, const double *lo, const double *hi,
                             const double *Tt, const double *init, double *q_out, double *x_raw, double *fnorm,
                             signed char *status, int *evals, size_t n)
{
    ChainConst<6> ch;
    fold_chain<6>(screws, home, ch);
    double blo[6], bhi[6];
    for (int i = 0; i < 6; i++)
        blo[i] = lo[i], bhi[i] = hi[i];
    for (size_t idx = 0; idx < n; idx++)
    {
        CodoLane st;
        double mem[CODO_LANE_DOUBLES];
        codo_attach(st, LaneArray{mem});
        double m[16], q0[6];
        std::memcpy(m, Tt + 16 * idx, sizeof(m));
        for (int k = 0; k < 6; k++)
            st.x[k] = q0[k] = init[6 * idx + k];
        Rigid pose, poseI;
        rigid_from_record(m, pose);
        rigid_inv(pose, poseI);
        bool finished = codo_begin(blo, bhi, st);
        int count = 0;
        while (!finished)
        {
            double fp[6];
            codo_eval<6>(ch, poseI, st.probe, fp, codo_jac_probe(st));
            count++;
            finished = codo_advance(1u, blo, bhi, st, fp);
        }
        for (int k = 0; k < 6; k++)
        {
            q_out[6 * idx + k] = st.stat == PPX_STATUS_CONVERGED ? st.x[k] : q0[k];
            if (x_raw)
                x_raw[6 * idx + k] = st.x[k];
        }
        if (fnorm)
            fnorm[idx] = st.y;
        if (status)
            status[idx] = (signed char)st.stat;
        if (evals)
            evals[idx] = count;
    }
    return 0;
}
