// host_api_test.cpp -- the reference's own test bodies for the hot path (test/lie_test.cpp, test/matrix_test.cpp),
// replayed against include/ppx/ppx.hpp: same names, same literals, same EPS_SP comparison (Matrix::operator==).
// Every numerical routine below runs on the GPU through libppx_cuda.so.  Built and run by
// tests/test_gpu_parity.py::test_cpp_host_api (C++14, g++).
#include <cstdio>
#include <vector>

#include "ppx/ppx.hpp"

using namespace ppx;

static int failures = 0;
#define EXPECT(cond)                                                      \
    do                                                                    \
    {                                                                     \
        if (!(cond))                                                      \
        {                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                   \
        }                                                                 \
    } while (0)

constexpr double DEG_RAD(double deg) { return deg * PI / 180.0; }

int main()
{
    // test/lie_test.cpp:136-142
    {
        Matrix<3, 1> vec{1, 2, 3};
        SO3 expected{0, 3, -2, -3, 0, 1, 2, -1, 0};
        SO3 result = hat(vec);
        EXPECT(result == expected);
    }
    // test/lie_test.cpp:144-154
    for (int i = -179; i < 180; i += 7)
    {
        so3 w{0.0, 0.0, DEG_RAD((double)i)};
        auto R = w.exp();
        EXPECT(SO3::RotZ(DEG_RAD((double)i)) == R);
        auto dR = R.log();
        EXPECT(norm2(dR - w) < EPS_SP);
    }
    // test/lie_test.cpp:156-168
    {
        SE3 Tinput{1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 3, 1};
        Matrix<4, 4> expected{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.57079633, 0.0, 0.0, -1.57079633, 0.0, 0.0, 0.0, 2.35619449, 2.35619449, 0.0};
        auto result = hat(Tinput.log());
        EXPECT(result == expected);
    }
    // test/lie_test.cpp:170-182
    {
        Matrix<4, 4> se3mat = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, PI / 2.0, 0.0, 0.0, -PI / 2.0, 0.0, 0.0, 0.0, 2.3562, 2.3562, 0.0};
        SE3 result{SO3{1, 0, 0, 0, 0, 1, 0, -1, 0}, T3{0, 0, 3}};
        auto cal = vee(se3mat).exp();
        for (std::size_t i = 0; i < 16; i++)
            EXPECT(std::fabs(cal[i] - result[i]) < 1.0e-5);
    }
    // test/lie_test.cpp:184-216
    kinematics<6> UR5;
    SE3 F6{-1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.817, 0.191, -0.006, 1.0};
    UR5.setJoint<0>({"R1", se3{0, 0, 1, 0, 0, 0, 0}, SE3()});
    UR5.setJoint<1>({"R2", se3{0.0, 1.0, 0.0, -0.089, 0.0, 0.0}, SE3{}});
    UR5.setJoint<2>({"R3", se3{0.0, 1.0, 0.0, -0.089, 0.0, 0.425}, SE3{}});
    UR5.setJoint<3>({"R4", se3{0.0, 1.0, 0.0, -0.089, 0.0, 0.817}, SE3{}});
    UR5.setJoint<4>({"R5", se3{0.0, 0.0, -1.0, -0.109, 0.817, 0.0}, SE3{}});
    UR5.setJoint<5>({"R6", se3{0.0, 1.0, 0.0, 0.006, 0.0, 0.817}, F6});
    {
        SE3 result{0.0, 1.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.095, 0.109, 0.988, 1.0};
        kinematics<6>::Q q{0.0, -0.5 * PI, 0.0, 0.0, 0.5 * PI, 0.0};
        auto r = UR5.forwardSpace(q);
        EXPECT(r == result);
        // column 0 = S_0 as in demo/robotics.hpp:97 (the test's private copy leaves it zero)
        Matrix<6, 6> result2{0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, -0.089, 0.0, 0.0, 0.0, 1.0, 0.0, -0.514, 0.0, 0.0,
                             0.0, 1.0, 0.0, -0.906, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.906, -0.109, 0.0, 0.0, 1.0, 0.109, -0.095, 0.0};
        auto r2 = UR5.jacobiSpace(q);
        EXPECT(r2 == result2);
        SE3 TargetPose{0.0, 1.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.095, 0.109, 0.988, 1.0};
        auto j = UR5.inverseSpaceNewton(TargetPose, {0.0, -1.5, 0.0, 0.0, 1.5, 0.0});
        EXPECT(UR5.forwardSpace(j) == TargetPose);
        // demo/main.cpp:369-375: CoDo trust-region inverseSpace with (-pi, pi) joint ranges
        for (auto &jt : UR5.JointList())
            jt.range = {-PI, PI};
        kinematics<6>::Q qd{0.336058, 1.69911, -0.445658, -1.25227, 0.00013744, -2.71554};
        SE3 Td = UR5.forwardSpace(qd);
        auto qs = UR5.inverseSpace(Td, qd - 0.05);
        EXPECT(norm2(qs - qd) < 1e-5);
    }
    // test/matrix_test.cpp:819-850 with the documented spellings of README.md:70-95
    {
        Matrix<4, 3> X{3.5, 1.6, 3.7, 4.3, 2.7, -5.7, -0.8, -9.8, -3.1, -6.0, 1.9, 6.9};
        Matrix<4, 1> Y{1, 1, 1, 1};
        Matrix<3, 1> result{0.271349846985587, -0.030388140654028, -0.062228084118990};
        auto x1 = linsolve<Factorization::QR>(X, Y);
        auto x2 = linsolve<factorization::SVD>(X, Y);
        EXPECT(x1.x == result);
        EXPECT(x2.x == result);
        EXPECT(x1.s == StatusCode::NORMAL);
        Matrix<4, 3> u{1, 4, 7, 11, 2, 5, 8, 1, 3, 6, 9, 5};
        SVD<4, 3> ru(u);
        EXPECT(ru.u * ru.w.diag() * ru.v.T() == u);
        Matrix<3, 1> w;
        Matrix<3, 3> v;
        bool sing = true;
        auto uu = svdcmp(u, w, v, sing);
        EXPECT(uu * w.diag() * v.T() == u && !sing);
        Matrix<4, 4> g{2, -1, 1, 4, -1, 2, -1, 1, 1, -1, 2, -2, 4, 1, -2, 3};
        EigenValue<4> e1(g, true);
        Matrix<4, 1> result2{-2.674416988218162, 1.0, 3.945104914279287, 6.729312073938871};
        EXPECT(e1.d == result2);
        EXPECT(eig<eigensystem::SymOnlyVal>(g) == result2);
        auto ev = eig<eigensystem::SymValAndVecSorted>(g);
        EXPECT(g * ev.vec == ev.vec * ev.val.diag()); // eigenvectors are defined up to sign: check S z = z d
    }
    // test/matrix_test.cpp:809-817 and test/solver_test.cpp:86-103
    {
        Matrix<2, 2> x = {1, 2, 4, 3};
        Matrix<2, 2> result{-0.6, 0.4, 0.8, -0.2};
        EXPECT(std::fabs(x.det() - (-5.0)) < 1e-14);
        EXPECT(x.I() == result);
        EXPECT(inverse(x) == result && std::fabs(determinant(x) + 5.0) < 1e-14);
        Matrix<4, 4> C{0.438744359656398, 0.381558457093008, 0.765516788149002, 0.795199901137063,
                       0.186872604554379, 0.489764395788231, 0.445586200710900, 0.646313010111265,
                       0.709364830858073, 0.754686681982361, 0.276025076998578, 0.679702676853675,
                       0.655098003973841, 0.162611735194631, 0.118997681558377, 0.498364051982143};
        EXPECT(std::fabs(norm2(C) - 2.075457904661895) < 10 * EPS_DP);
        Matrix<4, 1> rhs{1, 2, 3, 4};
        auto lu = linsolve<Factorization::LU>(C, rhs);
        EXPECT(lu.s == StatusCode::NORMAL && C * lu.x == rhs);
        EXPECT(C * pinv(C) == eye<4>());
    }
    // the batched entry points on host vectors of reference records
    {
        const std::size_t n = 100000;
        std::vector<kinematics<6>::Q> q(n);
        for (std::size_t i = 0; i < n; i++)
            for (int k = 0; k < 6; k++)
                q[i][k] = std::sin(0.001 * (double)(i * 6 + k)) * PI;
        std::vector<SE3> T(n);
        std::vector<Matrix<6, 6>> Js(n);
        cuda::batch_fk_jacobian(UR5, q.data(), T.data(), Js.data(), n);
        EXPECT(T[777] == UR5.forwardSpace(q[777]));
        EXPECT(Js[99999] == UR5.jacobiSpace(q[99999]));
        std::vector<se3> xi(n), back(n);
        for (std::size_t i = 0; i < n; i++)
            xi[i] = se3{0.3 * std::sin(i), 0.2, -0.4 * std::cos(i), 1.0, -2.0, 0.5};
        cuda::batch_exp_log(xi.data(), back.data(), n);
        EXPECT(back[4242] == xi[4242]);
        std::vector<Matrix<4, 3>> A(n);
        std::vector<Matrix<4, 1>> b(n);
        for (std::size_t i = 0; i < n; i++)
        {
            for (int k = 0; k < 12; k++)
            {
                const double h = std::sin(12.9898 * (double)i + 78.233 * (double)k) * 43758.5453; // hash -> (-1, 1)
                A[i][k] = 2.0 * (h - std::floor(h)) - 1.0;
            }
            b[i] = {1, 1, 1, 1};
        }
        std::vector<EqnResult<3>> res(n);
        cuda::batch_linsolve<factorization::QR>(A.data(), b.data(), res.data(), n);
        EXPECT(res[31337].x == linsolve<factorization::SVD>(A[31337], b[31337]).x);
        // MatrixS::operator* (include/matrixs.hpp:596-640) batched: exact k-j-i accumulation, checked bit for bit
        // against the same loop written out here (this file is compiled without FMA contraction for that)
        std::vector<Matrix<3, 3>> B3(n);
        std::vector<Matrix<4, 3>> AB(n);
        for (std::size_t i = 0; i < n; i++)
            for (int k = 0; k < 9; k++)
                B3[i][k] = std::cos(0.37 * (double)(i + 1) * (k + 1));
        cuda::batch_mul(A.data(), B3.data(), AB.data(), n); // (4,3,3): the generic kernel
        bool same = true;
        for (std::size_t i = 0; i < n; i += 997)
        {
            double c[12] = {};
            for (int k = 0; k < 3; k++)
                for (int j = 0; j < 3; j++)
                    for (int r = 0; r < 4; r++)
                        c[r + j * 4] += A[i][r + k * 4] * B3[i][k + j * 3];
            for (int e = 0; e < 12; e++)
                same = same && c[e] == AB[i][e];
        }
        EXPECT(same);
        // the whole box behind the same call: every visible device, then back to the current one
        cuda::use_all_devices();
        std::vector<SE3> T2(n);
        std::vector<Matrix<6, 6>> Js2(n);
        cuda::batch_fk_jacobian(UR5, q.data(), T2.data(), Js2.data(), n);
        cuda::use_current_device();
        const std::vector<int> tuned = cuda::tune_devices(1, 0.05); // measured choice of devices: never empty
        EXPECT(!tuned.empty());
        cuda::use_current_device();
        bool bitwise = true;
        for (std::size_t i = 0; i < n; i++)
            bitwise = bitwise && std::equal(T[i].begin(), T[i].end(), T2[i].begin()) &&
                      std::equal(Js[i].begin(), Js[i].end(), Js2[i].begin());
        EXPECT(bitwise);
    }
    // constructors of include/liegroup.hpp:41-43, 200
    {
        SO3 R(so3{0, 1, 0}, so3{-1, 0, 0}, so3{0, 0, 1});
        EXPECT(R == SO3::RotZ(PI / 2));
        double rm[3][3] = {{0, -1, 0}, {1, 0, 0}, {0, 0, 1}};
        EXPECT(SO3(rm) == R);
        double tm[4][4] = {{0, -1, 0, 0.5}, {1, 0, 0, -2}, {0, 0, 1, 3}, {9, 9, 9, 9}};
        SE3 Tm(tm);
        EXPECT(Tm == SE3(R, T3{0.5, -2.0, 3.0}) && Tm(3, 0) == 0.0 && Tm(3, 3) == 1.0);
        EXPECT((Tm * Tm.I()) == SE3());
    }
    std::printf(failures ? "host_api_test: %d FAILURES\n" : "host_api_test: all passed\n", failures);
    return failures ? 1 : 0;
}
