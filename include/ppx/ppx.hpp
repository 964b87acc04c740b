// ppx.hpp -- header-only C++14 host API over libppx_cuda.so (include/ppx_cuda.h).
//
// Keeps the call surface of the reference library (Xtinc/matrix, "ppx") for its data-parallel hot path --
// Matrix / MatrixS, SO3, SE3, so3, se3, hat/vee, linsolve<>, SVD / svdcmp, EigenValue / eig<>, kinematics<N> --
// under both the names the reference's headers use (MatrixS, Factorization, SVD<m,n>, EigenValue<n>) and the
// names its README / manual document (Matrix, factorization, svdcmp, eig<eigensystem::...>), and adds the
// batched entry points ppx::cuda::batch_* that evaluate millions of independent instances in one launch.
//
// What lives where
//   - Containers (Matrix<M,N>: column-major doubles, sizeof == M*N*8, list constructors that truncate /
//     zero-pad like the reference's, element access, transpose, +, -, scalar products, ==) are plain host
//     code: they carry data, they are not the hot path.
//   - Every hot-path ROUTINE (matrix product, exp, log, Adt, I, SE3 product, linsolve, SVD, EigenValue,
//     forwardSpace, jacobiSpace, inverseSpace) runs on the GPU.  The batch_* functions take host arrays of reference
//     records -- std::vector<ppx::Matrix<M,N>>::data() -- and call the ppx_cuda_host_* entry points, which
//     overlap H2D / kernel / D2H.  The single-value members (xi.exp(), T.log(), linsolve<>(A, b), ...)
//     exist for source compatibility and are batches of one: there is no CPU implementation of the math
//     in this header, so without a CUDA device they throw ppx::cuda::Error.
//
// Reference citations (file:line) are relative to the reference tree.
#ifndef PPX_B200_HOST_API_HPP
#define PPX_B200_HOST_API_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../ppx_cuda.h"

namespace ppx
{
constexpr double PI = 3.141592653589793;                              // include/exprtmpl.hpp:11
constexpr double EPS_SP = std::numeric_limits<float>::epsilon();       // :12
constexpr double EPS_DP = std::numeric_limits<double>::epsilon();      // :13
constexpr double MAX_SP = std::numeric_limits<float>::max();           // :14

namespace cuda
{
struct Error : std::runtime_error
{
    int code;
    explicit Error(int c) : std::runtime_error(std::string("libppx_cuda: ") + ppx_cuda_error_string(c)), code(c) {}
};
inline void check(int code)
{
    if (code != PPX_OK)
        throw Error(code);
}
} // namespace cuda

class SO3;
class SE3;

// ---------------------------------------------------------------------------------------------------------
// Matrix<M,N> -- the reference's MatrixS<M,N> (include/matrixs.hpp:80-81): column-major, element (r,c) at
// data()[r + c*M] (:572-582), value-initialised to zero, no padding (sizeof(Matrix<2,2>) == 32,
// test/matrix_test.cpp:773).
template <std::size_t M, std::size_t N>
class Matrix
{
public:
    static constexpr std::size_t ROW = M, COL = N, LEN = M * N;

    Matrix() : m_{} {}
    // list constructor: fills column-major, silently truncates / zero-pads (include/matrixs.hpp:380-399)
    template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    Matrix(std::initializer_list<T> list) : m_{}
    {
        std::size_t i = 0;
        for (auto it = list.begin(); it != list.end() && i < LEN; ++it, ++i)
            m_[i] = static_cast<double>(*it);
    }
    // (omega, v) constructor of se3 (include/matrixs.hpp:744-754)
    template <std::size_t A = M, std::size_t B = N, typename = typename std::enable_if<A == 6 && B == 1>::type>
    Matrix(const Matrix<3, 1> &w, const Matrix<3, 1> &v) : m_{}
    {
        for (int i = 0; i < 3; i++)
        {
            m_[i] = w[i];
            m_[3 + i] = v[i];
        }
    }

    double *data() { return m_.data(); }
    const double *data() const { return m_.data(); }
    double &operator()(std::size_t r, std::size_t c) { return m_.at(r + c * M); }
    const double &operator()(std::size_t r, std::size_t c) const { return m_.at(r + c * M); }
    double &operator[](std::size_t i) { return m_.at(i); }
    const double &operator[](std::size_t i) const { return m_.at(i); }
    double *begin() { return m_.data(); }
    double *end() { return m_.data() + LEN; }
    const double *begin() const { return m_.data(); }
    const double *end() const { return m_.data() + LEN; }
    const double *cbegin() const { return m_.data(); }
    const double *cend() const { return m_.data() + LEN; }
    void fill(double v) { m_.fill(v); }

    Matrix<N, M> T() const
    {
        Matrix<N, M> r;
        for (std::size_t i = 0; i < M; i++)
            for (std::size_t j = 0; j < N; j++)
                r(j, i) = (*this)(i, j);
        return r;
    }
    double trace() const
    {
        static_assert(M == N, "only square matrix has a trace.");
        double t = 0.0;
        for (std::size_t i = 0; i < M; i++)
            t += (*this)(i, i);
        return t;
    }
    // vector -> diagonal matrix, matrix -> diagonal vector (include/matrixs.hpp:549-569)
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<(A == 1 || B == 1), Matrix<(A > B ? A : B), (A > B ? A : B)>>::type diag() const
    {
        Matrix<(A > B ? A : B), (A > B ? A : B)> r;
        for (std::size_t i = 0; i < LEN; i++)
            r(i, i) = m_[i];
        return r;
    }
    // MatrixS::operator*, include/matrixs.hpp:596-640: a batch of one through ppx_cuda_host_matmul (k-j-i
    // accumulation, products and sums rounded separately like the reference's loop); dimensions up to 8
    template <std::size_t L>
    Matrix<M, L> operator*(const Matrix<N, L> &o) const
    {
        static_assert(M <= 8 && N <= 8 && L <= 8, "libppx_cuda multiplies matrices with dimensions up to 8");
        Matrix<M, L> r;
        cuda::check(ppx_cuda_host_matmul((int)M, (int)N, (int)L, data(), o.data(), r.data(), 1));
        return r;
    }
    Matrix operator+(const Matrix &o) const { return zip(o, [](double a, double b) { return a + b; }); }
    Matrix operator-(const Matrix &o) const { return zip(o, [](double a, double b) { return a - b; }); }
    Matrix &operator+=(const Matrix &o) { return *this = *this + o; }
    Matrix &operator-=(const Matrix &o) { return *this = *this - o; }
    Matrix operator*(double s) const { return map([s](double a) { return a * s; }); }
    Matrix operator/(double s) const { return map([s](double a) { return a / s; }); }
    Matrix operator+(double s) const { return map([s](double a) { return a + s; }); }
    Matrix operator-(double s) const { return map([s](double a) { return a - s; }); }
    friend Matrix operator*(double s, const Matrix &a) { return a.map([s](double x) { return s * x; }); }
    // element-wise |a-b| < EPS_SP, the comparison every EXPECT_EQ of the reference's tests uses (:660-665)
    bool operator==(const Matrix &o) const
    {
        for (std::size_t i = 0; i < LEN; i++)
            if (!(std::fabs(m_[i] - o.m_[i]) < EPS_SP))
                return false;
        return true;
    }
    bool operator!=(const Matrix &o) const { return !(*this == o); }
    template <std::size_t A, std::size_t B>
    Matrix<A, B> sub(std::size_t r0, std::size_t c0) const // read-only block (include/matrixs.hpp:471-490)
    {
        Matrix<A, B> r;
        for (std::size_t j = 0; j < B; j++)
            for (std::size_t i = 0; i < A; i++)
                r(i, j) = (*this)(r0 + i, c0 + j);
        return r;
    }
    // inverse / determinant through LU on the GPU (include/linalg.hpp:933-965); defined after the class
    Matrix I() const;
    double det() const;
    // se3 halves (include/matrixs.hpp:732-742)
    template <std::size_t A = M, std::size_t B = N, typename = typename std::enable_if<A == 6 && B == 1>::type>
    Matrix<3, 1> _1() const { return {m_[0], m_[1], m_[2]}; }
    template <std::size_t A = M, std::size_t B = N, typename = typename std::enable_if<A == 6 && B == 1>::type>
    Matrix<3, 1> _2() const { return {m_[3], m_[4], m_[5]}; }

    // Lie-algebra members of 3x1 / 6x1 vectors (declared include/matrixs.hpp:707-730); defined below, GPU-backed
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, SO3>::type exp() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 6 && B == 1, SE3>::type exp() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type adt() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 6 && B == 1, Matrix<6, 6>>::type adt() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type ljac() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type ljacinv() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type rjac() const;
    template <std::size_t A = M, std::size_t B = N>
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type rjacinv() const;

private:
    template <class F>
    Matrix map(F f) const
    {
        Matrix r;
        for (std::size_t i = 0; i < LEN; i++)
            r.m_[i] = f(m_[i]);
        return r;
    }
    template <class F>
    Matrix zip(const Matrix &o, F f) const
    {
        Matrix r;
        for (std::size_t i = 0; i < LEN; i++)
            r.m_[i] = f(m_[i], o.m_[i]);
        return r;
    }
    std::array<double, M * N> m_;
};

template <std::size_t M, std::size_t N>
Matrix<M, N> Matrix<M, N>::I() const
{
    static_assert(M == N, "only square matrix has an inverse. For rectangular matrix ,query for pinv.");
    Matrix r;
    cuda::check(ppx_cuda_host_inverse((int)M, data(), r.data(), 1));
    return r;
}
template <std::size_t M, std::size_t N>
double Matrix<M, N>::det() const
{
    static_assert(M == N, "only square matrix has a determinant.");
    double d = 0.0;
    cuda::check(ppx_cuda_host_det((int)M, data(), &d, 1));
    return d;
}

template <std::size_t M, std::size_t N>
using MatrixS = Matrix<M, N>; // the name the reference's headers use

template <std::size_t N>
Matrix<N, N> eye() // include/linalg.hpp:112-121
{
    Matrix<N, N> r;
    for (std::size_t i = 0; i < N; i++)
        r(i, i) = 1.0;
    return r;
}

using so3 = Matrix<3, 1>; // include/liegroup.hpp:8-10
using se3 = Matrix<6, 1>;
using T3 = Matrix<3, 1>;

inline Matrix<3, 3> hat(const so3 &v) { return {0.0, v[2], -v[1], -v[2], 0.0, v[0], v[1], -v[0], 0.0}; } // :12-15
inline so3 vee(const Matrix<3, 3> &m) { return {m[5], m[6], m[1]}; }                                      // :17-20
inline Matrix<4, 4> hat(const se3 &v)                                                                     // :22-25
{
    return {0.0, v[2], -v[1], 0.0, -v[2], 0.0, v[0], 0.0, v[1], -v[0], 0.0, 0.0, v[3], v[4], v[5], 0.0};
}
inline se3 vee(const Matrix<4, 4> &m) { return {m[6], m[8], m[1], m[12], m[13], m[14]}; }                 // :27-30

// include/liegroup.hpp:32-114
class SO3 : public Matrix<3, 3>
{
public:
    SO3() : Matrix<3, 3>(eye<3>()) {}
    SO3(const Matrix<3, 3> &o) : Matrix<3, 3>(o) {}
    template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    SO3(std::initializer_list<T> l) : Matrix<3, 3>(l) {}
    SO3(const Matrix<3, 1> &xAxis, const Matrix<3, 1> &yAxis, const Matrix<3, 1> &zAxis) // :41-42: the axes are the columns
        : Matrix<3, 3>{xAxis[0], xAxis[1], xAxis[2], yAxis[0], yAxis[1], yAxis[2], zAxis[0], zAxis[1], zAxis[2]}
    {
    }
    SO3(double m[3][3]) // :43: row-major C array
        : Matrix<3, 3>{m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]}
    {
    }
    using Matrix<3, 3>::operator*;
    SO3 operator*(const SO3 &o) const { return SO3(Matrix<3, 3>::operator*(o)); }
    Matrix<3, 3> Adt() const { return *this; }
    SO3 I() const { return SO3(this->T()); }
    so3 log() const
    {
        so3 w;
        cuda::check(ppx_cuda_host_SO3_log(data(), w.data(), 1));
        return w;
    }
    static SO3 RotX(double t) { return {1.0, 0.0, 0.0, 0.0, std::cos(t), -std::sin(t), 0.0, std::sin(t), std::cos(t)}; }
    static SO3 RotY(double t) { return {std::cos(t), 0.0, -std::sin(t), 0.0, 1.0, 0.0, std::sin(t), 0.0, std::cos(t)}; }
    static SO3 RotZ(double t) { return {std::cos(t), std::sin(t), 0.0, -std::sin(t), std::cos(t), 0.0, 0.0, 0.0, 1.0}; }
};

// include/liegroup.hpp:186-251
class SE3 : public Matrix<4, 4>
{
public:
    SE3() : Matrix<4, 4>(eye<4>()) {}
    SE3(const Matrix<4, 4> &o) : Matrix<4, 4>(o) {}
    template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    SE3(std::initializer_list<T> l) : Matrix<4, 4>(l) {}
    SE3(const SO3 &R, const T3 &p) : Matrix<4, 4>(eye<4>())
    {
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
                (*this)(r, c) = R(r, c);
        for (int r = 0; r < 3; r++)
            (*this)(r, 3) = p[r];
    }
    SE3(double m[4][4]) // :200: row-major C array; the bottom row is set to (0, 0, 0, 1) whatever m holds
        : Matrix<4, 4>{m[0][0], m[1][0], m[2][0], 0.0, m[0][1], m[1][1], m[2][1], 0.0,
                       m[0][2], m[1][2], m[2][2], 0.0, m[0][3], m[1][3], m[2][3], 1.0}
    {
    }
    using Matrix<4, 4>::operator*;
    SO3 Rot() const { return SO3(sub<3, 3>(0, 0)); }
    T3 Pos() const { return sub<3, 1>(0, 3); }
    SE3 operator*(const SE3 &o) const
    {
        SE3 r;
        cuda::check(ppx_cuda_host_SE3_mul(data(), o.data(), r.data(), 1));
        return r;
    }
    Matrix<6, 6> Adt() const
    {
        Matrix<6, 6> r;
        cuda::check(ppx_cuda_host_SE3_Adt(data(), r.data(), 1));
        return r;
    }
    SE3 I() const
    {
        SE3 r;
        cuda::check(ppx_cuda_host_SE3_inv(data(), r.data(), 1));
        return r;
    }
    se3 log() const
    {
        se3 r;
        cuda::check(ppx_cuda_host_SE3_log(data(), r.data(), 1));
        return r;
    }
};

static_assert(sizeof(Matrix<2, 2>) == 32 && sizeof(se3) == 48 && sizeof(SO3) == 72 && sizeof(SE3) == 128 &&
                  sizeof(Matrix<6, 6>) == 288 && sizeof(Matrix<4, 3>) == 96,
              "Matrix<M,N> must be a dense record: the batch ABI reinterprets arrays of them as double*");

template <std::size_t M, std::size_t N>
template <std::size_t A, std::size_t B>
typename std::enable_if<A == 3 && B == 1, SO3>::type Matrix<M, N>::exp() const // include/liegroup.hpp:116-133
{
    SO3 r;
    cuda::check(ppx_cuda_host_so3_exp(data(), r.data(), 1));
    return r;
}
template <std::size_t M, std::size_t N>
template <std::size_t A, std::size_t B>
typename std::enable_if<A == 6 && B == 1, SE3>::type Matrix<M, N>::exp() const // :253-269
{
    SE3 r;
    cuda::check(ppx_cuda_host_se3_exp(data(), r.data(), 1));
    return r;
}
template <std::size_t M, std::size_t N>
template <std::size_t A, std::size_t B>
typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type Matrix<M, N>::adt() const { return hat(*this); } // :135-140
template <std::size_t M, std::size_t N>
template <std::size_t A, std::size_t B>
typename std::enable_if<A == 6 && B == 1, Matrix<6, 6>>::type Matrix<M, N>::adt() const // :271-280
{
    Matrix<6, 6> r;
    cuda::check(ppx_cuda_host_se3_adt(data(), r.data(), 1));
    return r;
}
#define PPX_SO3_JAC(name, which)                                                                   \
    template <std::size_t M, std::size_t N>                                                        \
    template <std::size_t A, std::size_t B>                                                        \
    typename std::enable_if<A == 3 && B == 1, Matrix<3, 3>>::type Matrix<M, N>::name() const       \
    {                                                                                              \
        Matrix<3, 3> r;                                                                            \
        cuda::check(ppx_cuda_host_so3_jac(which, data(), r.data(), 1));                            \
        return r;                                                                                  \
    }
PPX_SO3_JAC(ljac, 0)    // include/liegroup.hpp:142-155
PPX_SO3_JAC(ljacinv, 1) // :157-170
PPX_SO3_JAC(rjac, 2)    // :172-177
PPX_SO3_JAC(rjacinv, 3) // :179-184
#undef PPX_SO3_JAC

// ---------------------------------------------------------------------------------------------------------
// include/linalg.hpp:362-404
enum class StatusCode : char
{
    NORMAL,
    CONVERGED,
    DIVERGED,
    SINGULAR
};
template <std::size_t N>
struct EqnResult
{
    Matrix<N, 1> x;
    StatusCode s = StatusCode::NORMAL;
};
enum class Factorization : char // include/linalg.hpp:1169-1174
{
    LU,
    QR,
    SVD
};
using factorization = Factorization; // the spelling of README.md:70-71
enum class eigensystem : char         // README.md:94-95
{
    SymOnlyVal,
    SymValAndVec,
    SymValAndVecSorted
};

// SVD<m,n>: constructor decomposes, members u, w, v with A = u diag(w) v^T (include/linalg.hpp:605-924)
template <int m, int n>
class SVD
{
public:
    explicit SVD(const Matrix<m, n> &a)
    {
        cuda::check(ppx_cuda_host_svd(m, n, a.data(), u.data(), w.data(), v.data(), 1));
    }
    void solve(double *b) const // :886-913, same truncation; b holds max(m, n) entries, the first n are x
    {
        Matrix<m, n> a = u * w.diag() * v.T();
        Matrix<n, 1> x;
        cuda::check(ppx_cuda_host_linsolve_svd(m, n, a.data(), b, x.data(), 1));
        std::copy(x.begin(), x.end(), b);
    }
    double norm() const { return *std::max_element(w.cbegin(), w.cend()); } // :915-918
    Matrix<m, n> u;
    Matrix<n, 1> w;
    Matrix<n, n> v;
};

// documented free function (docs/ppx Reference Manual.md:499-520, README.md:78-86)
template <std::size_t M, std::size_t N>
Matrix<M, N> svdcmp(Matrix<M, N> u, Matrix<N, 1> &w, Matrix<N, N> &v)
{
    Matrix<M, N> a = u;
    cuda::check(ppx_cuda_host_svd((int)M, (int)N, a.data(), u.data(), w.data(), v.data(), 1));
    return u;
}
template <std::size_t M, std::size_t N>
Matrix<M, N> svdcmp(Matrix<M, N> u, Matrix<N, 1> &w, Matrix<N, N> &v, bool &sing)
{
    sing = false;
    return svdcmp(u, w, v);
}

// EigenValue<n>(S, sorted): members d (values), z (vectors in columns) (include/linalg.hpp:967-1167)
template <int n>
class EigenValue
{
public:
    explicit EigenValue(const Matrix<n, n> &s, bool sorted = false)
    {
        cuda::check(ppx_cuda_host_eig_sym(n, s.data(), d.data(), z.data(), sorted ? 1 : 0, 1));
    }
    Matrix<n, n> z;
    Matrix<n, 1> d;
};
template <std::size_t N>
struct EigResult
{
    Matrix<N, N> vec;
    Matrix<N, 1> val;
};
namespace details
{
template <eigensystem type, std::size_t N>
struct eig_impl
{
    static EigResult<N> run(const Matrix<N, N> &s)
    {
        EigResult<N> r;
        cuda::check(ppx_cuda_host_eig_sym((int)N, s.data(), r.val.data(), r.vec.data(),
                                          type == eigensystem::SymValAndVecSorted ? 1 : 0, 1));
        return r;
    }
};
template <std::size_t N>
struct eig_impl<eigensystem::SymOnlyVal, N>
{
    static Matrix<N, 1> run(const Matrix<N, N> &s)
    {
        Matrix<N, 1> d;
        cuda::check(ppx_cuda_host_eig_sym((int)N, s.data(), d.data(), nullptr, 1, 1));
        return d;
    }
};
} // namespace details
template <eigensystem type, std::size_t N>
auto eig(const Matrix<N, N> &s) -> decltype(details::eig_impl<type, N>::run(s))
{
    return details::eig_impl<type, N>::run(s);
}

// MatrixS::I(), det(), pinv(), matrix norm2 (include/linalg.hpp:933-965, 1211-1227, 926-931); the reference spells the
// first two as members, its README as free functions inverse() / determinant()
template <std::size_t N>
Matrix<N, N> inverse(const Matrix<N, N> &a)
{
    Matrix<N, N> r;
    cuda::check(ppx_cuda_host_inverse((int)N, a.data(), r.data(), 1));
    return r;
}
template <std::size_t N>
double determinant(const Matrix<N, N> &a)
{
    double d = 0.0;
    cuda::check(ppx_cuda_host_det((int)N, a.data(), &d, 1));
    return d;
}
template <std::size_t M, std::size_t N>
Matrix<N, M> pinv(const Matrix<M, N> &a)
{
    Matrix<N, M> r;
    cuda::check(ppx_cuda_host_pinv((int)M, (int)N, a.data(), r.data(), 1));
    return r;
}
// vector norm2 = sqrt(inner_product) (include/linalg.hpp:217-241): a container-level helper, host code
template <std::size_t M, std::size_t N>
typename std::enable_if<(M == 1 || N == 1), double>::type norm2(const Matrix<M, N> &a)
{
    double s = 0.0;
    for (double x : a)
        s += x * x;
    return std::sqrt(s);
}
template <std::size_t M, std::size_t N>
typename std::enable_if<(M > 1 && N > 1), double>::type norm2(const Matrix<M, N> &a)
{
    double s = 0.0;
    cuda::check(ppx_cuda_host_norm2((int)M, (int)N, a.data(), &s, 1));
    return s;
}

// linsolve<Factorization::LU|QR|SVD>(A, b) (include/linalg.hpp:1176-1209)
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::LU, EqnResult<N>>::type linsolve(const Matrix<M, N> &A, Matrix<M, 1> b)
{
    static_assert(M == N, "LU decomposation only supports square matrix.");
    EqnResult<N> r;
    int8_t st = 0;
    cuda::check(ppx_cuda_host_linsolve_lu((int)N, A.data(), b.data(), r.x.data(), &st, 1));
    r.s = static_cast<StatusCode>(st);
    return r;
}
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::QR, EqnResult<N>>::type linsolve(const Matrix<M, N> &A, Matrix<M, 1> b)
{
    static_assert(M >= N, "QR decomposation only supports column full-rank matrix.");
    EqnResult<N> r;
    int8_t st = 0;
    cuda::check(ppx_cuda_host_linsolve_qr((int)M, (int)N, A.data(), b.data(), r.x.data(), &st, 1));
    r.s = static_cast<StatusCode>(st);
    return r;
}
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::SVD, EqnResult<N>>::type linsolve(const Matrix<M, N> &A, Matrix<M, 1> b)
{
    EqnResult<N> r;
    cuda::check(ppx_cuda_host_linsolve_svd((int)M, (int)N, A.data(), b.data(), r.x.data(), 1));
    return r; // always NORMAL, include/linalg.hpp:1208
}

// ---------------------------------------------------------------------------------------------------------
// demo/robotics.hpp
struct joint
{
    std::string name{"Anon"};
    std::pair<double, double> range{-MAX_SP, MAX_SP};
    se3 screw;
    SE3 pose;
    joint() = default;
    joint(const std::string &n, const se3 &s, const SE3 &p) : name(n), screw(s), pose(p) {}
    joint(const std::string &n, const se3 &s, const SE3 &p, double lo, double hi) : joint(n, s, p)
    {
        range.first = lo;
        range.second = hi;
    }
};

template <std::size_t N>
class kinematics
{
    std::array<joint, N> JList;

public:
    using Q = Matrix<N, 1>;
    template <std::size_t L>
    joint &Joint()
    {
        static_assert(L < N, "joint index out of range");
        return JList[L];
    }
    template <std::size_t L>
    const joint &Joint() const
    {
        static_assert(L < N, "joint index out of range");
        return JList[L];
    }
    template <std::size_t L>
    void setJoint(const joint &j) // spelling of test/lie_test.cpp:36-40
    {
        Joint<L>() = j;
    }
    std::array<joint, N> &JointList() { return JList; }
    const std::array<joint, N> &JointList() const { return JList; }

    // dense screw table (N records of 6) and home pose, the form the C ABI takes
    std::array<double, 6 * N> screws() const
    {
        std::array<double, 6 * N> s{};
        for (std::size_t i = 0; i < N; i++)
            std::copy(JList[i].screw.begin(), JList[i].screw.end(), s.begin() + 6 * i);
        return s;
    }
    const SE3 &home() const { return JList.back().pose; }

    SE3 forwardSpace(const Q &q) const // demo/robotics.hpp:83-92
    {
        SE3 T;
        auto s = screws();
        cuda::check(ppx_cuda_host_poe_fk_jacobian(s.data(), home().data(), (int)N, q.data(), T.data(), nullptr, 1));
        return T;
    }
    Matrix<6, N> jacobiSpace(const Q &q) const // :94-106
    {
        Matrix<6, N> J;
        auto s = screws();
        cuda::check(ppx_cuda_host_poe_fk_jacobian(s.data(), home().data(), (int)N, q.data(), nullptr, J.data(), 1));
        return J;
    }
    // bounded trust-region dogleg (details::CoDo) with the joints' ranges, demo/robotics.hpp:126-152:
    // the optimiser's result if it converged, else init
    Q inverseSpace(const SE3 &pose, const Q &init) const
    {
        static_assert(N == 6, "CoDo<N,6> needs a square Jacobian (linsolve<LU>, optimization.hpp:608)");
        Q q;
        auto s = screws();
        double lo[N], hi[N];
        for (std::size_t i = 0; i < N; i++)
        {
            lo[i] = JList[i].range.first;
            hi[i] = JList[i].range.second;
        }
        cuda::check(ppx_cuda_host_ik_codo(s.data(), home().data(), (int)N, lo, hi, pose.data(), init.data(), q.data(),
                                          nullptr, nullptr, nullptr, 1));
        return q;
    }
    // Newton-SVD loop of the test suite's private kinematics copy, test/lie_test.cpp:96-113
    Q inverseSpaceNewton(const SE3 &pose, Q init, int max_iter = 20) const
    {
        Q q;
        auto s = screws();
        cuda::check(ppx_cuda_host_ik_solve(s.data(), home().data(), (int)N, init.data(), pose.data(), max_iter,
                                           q.data(), nullptr, nullptr, 1));
        return q;
    }
};

// ---------------------------------------------------------------------------------------------------------
// Batched entry points: n independent instances per call, host arrays of reference records in and out.
namespace cuda
{
inline const double *dp(const void *p) { return static_cast<const double *>(p); }
inline double *dp(void *p) { return static_cast<double *>(p); }

inline void batch_exp(const so3 *w, SO3 *R, std::size_t n) { check(ppx_cuda_host_so3_exp(dp(w), dp(R), n)); }
inline void batch_log(const SO3 *R, so3 *w, std::size_t n) { check(ppx_cuda_host_SO3_log(dp(R), dp(w), n)); }
inline void batch_exp(const se3 *xi, SE3 *T, std::size_t n) { check(ppx_cuda_host_se3_exp(dp(xi), dp(T), n)); }
inline void batch_log(const SE3 *T, se3 *xi, std::size_t n) { check(ppx_cuda_host_SE3_log(dp(T), dp(xi), n)); }
// xi[i].exp().log() without materialising the SE3 (BASELINE config 1)
inline void batch_exp_log(const se3 *xi, se3 *out, std::size_t n) { check(ppx_cuda_host_se3_exp_log(dp(xi), dp(out), n)); }
inline void batch_mul(const SE3 *A, const SE3 *B, SE3 *C, std::size_t n) { check(ppx_cuda_host_SE3_mul(dp(A), dp(B), dp(C), n)); }
// MatrixS<M,N>::operator* over n pairs (include/matrixs.hpp:596-640), C[i] = A[i] * B[i]
template <std::size_t M, std::size_t N, std::size_t L>
void batch_mul(const Matrix<M, N> *A, const Matrix<N, L> *B, Matrix<M, L> *C, std::size_t n)
{
    static_assert(M <= 8 && N <= 8 && L <= 8, "libppx_cuda multiplies matrices with dimensions up to 8");
    check(ppx_cuda_host_matmul((int)M, (int)N, (int)L, dp(A), dp(B), dp(C), n));
}
// the devices the batch_* calls spread their instances over: use_devices({0, 1, 2, 3}), use_all_devices(), or
// use_current_device() (the default unless PPX_CUDA_DEVICES is set); see ppx_cuda_host_set_devices
inline void use_devices(const std::vector<int> &devices) { check(ppx_cuda_host_set_devices(devices.data(), (int)devices.size())); }
inline void use_all_devices() { check(ppx_cuda_host_set_devices(nullptr, -1)); }
inline void use_current_device() { check(ppx_cuda_host_set_devices(nullptr, 0)); }
// measures the host links of all visible devices (ppx_cuda_host_tune_devices, ~3 s for 8 devices) and makes the subset
// that carries the most together the device list; direction: 0 = calls dominated by inputs (H2D), 1 = by outputs (D2H),
// 2 = balanced.  -> the chosen ordinals
inline std::vector<int> tune_devices(int direction = 1, double seconds_per_probe = 0.15)
{
    const int have = ppx_cuda_device_count();
    if (have < 1)
        throw Error(PPX_ERR_NO_DEVICE);
    const std::size_t bytes = (std::size_t)have * (256u << 20);
    void *scratch = nullptr;
    check(ppx_cuda_malloc_host(&scratch, bytes));
    const int rc = ppx_cuda_host_tune_devices(nullptr, 0, scratch, bytes, direction, seconds_per_probe, nullptr);
    ppx_cuda_free_host(scratch);
    check(rc);
    std::vector<int> devs(have);
    const int count = ppx_cuda_host_get_devices(devs.data(), have);
    if (count < 0)
        throw Error(count);
    devs.resize((std::size_t)count);
    return devs;
}
inline void batch_I(const SE3 *T, SE3 *Ti, std::size_t n) { check(ppx_cuda_host_SE3_inv(dp(T), dp(Ti), n)); }
inline void batch_Adt(const SE3 *T, Matrix<6, 6> *Ad, std::size_t n) { check(ppx_cuda_host_SE3_Adt(dp(T), dp(Ad), n)); }
inline void batch_adt(const se3 *xi, Matrix<6, 6> *ad, std::size_t n) { check(ppx_cuda_host_se3_adt(dp(xi), dp(ad), n)); }
inline void batch_ljac(const so3 *w, Matrix<3, 3> *J, std::size_t n) { check(ppx_cuda_host_so3_jac(0, dp(w), dp(J), n)); }
inline void batch_ljacinv(const so3 *w, Matrix<3, 3> *J, std::size_t n) { check(ppx_cuda_host_so3_jac(1, dp(w), dp(J), n)); }
inline void batch_rjac(const so3 *w, Matrix<3, 3> *J, std::size_t n) { check(ppx_cuda_host_so3_jac(2, dp(w), dp(J), n)); }
inline void batch_rjacinv(const so3 *w, Matrix<3, 3> *J, std::size_t n) { check(ppx_cuda_host_so3_jac(3, dp(w), dp(J), n)); }

// kinematics<N>::forwardSpace / jacobiSpace for n joint vectors; T and/or Js may be null
template <std::size_t N>
void batch_fk_jacobian(const kinematics<N> &kin, const Matrix<N, 1> *q, SE3 *T, Matrix<6, N> *Js, std::size_t n)
{
    auto s = kin.screws();
    check(ppx_cuda_host_poe_fk_jacobian(s.data(), kin.home().data(), (int)N, dp(q), dp(T), dp(Js), n));
}
template <std::size_t N>
void batch_forwardSpace(const kinematics<N> &kin, const Matrix<N, 1> *q, SE3 *T, std::size_t n)
{
    batch_fk_jacobian<N>(kin, q, T, nullptr, n);
}
template <std::size_t N>
void batch_jacobiSpace(const kinematics<N> &kin, const Matrix<N, 1> *q, Matrix<6, N> *Js, std::size_t n)
{
    batch_fk_jacobian<N>(kin, q, nullptr, Js, n);
}
// one Newton step of test/lie_test.cpp:104-110 per instance; Vs may be null
template <std::size_t N>
void batch_ik_newton_step(const kinematics<N> &kin, const Matrix<N, 1> *q, const SE3 *target, Matrix<N, 1> *q_out,
                          se3 *Vs, std::size_t n)
{
    auto s = kin.screws();
    check(ppx_cuda_host_ik_newton_step(s.data(), kin.home().data(), (int)N, dp(q), dp(target), dp(q_out), dp(Vs), n));
}
// kinematics<6>::inverseSpace (CoDo trust region, joint ranges as the box) per instance; status may be null
template <std::size_t N>
void batch_inverseSpace(const kinematics<N> &kin, const SE3 *target, const Matrix<N, 1> *init, Matrix<N, 1> *q_out,
                        std::size_t n, StatusCode *status = nullptr)
{
    static_assert(N == 6, "CoDo<N,6> needs a square Jacobian");
    auto s = kin.screws();
    double lo[N], hi[N];
    for (std::size_t i = 0; i < N; i++)
    {
        lo[i] = kin.JointList()[i].range.first;
        hi[i] = kin.JointList()[i].range.second;
    }
    check(ppx_cuda_host_ik_codo(s.data(), kin.home().data(), (int)N, lo, hi, dp(target), dp(init), dp(q_out), nullptr,
                                nullptr, reinterpret_cast<int8_t *>(status), n));
}
// the whole Newton-SVD loop of test/lie_test.cpp:96-113 per instance; iters / status may be null
template <std::size_t N>
void batch_inverseSpaceNewton(const kinematics<N> &kin, const SE3 *target, const Matrix<N, 1> *init,
                              Matrix<N, 1> *q_out, std::size_t n, int max_iter = 20, int32_t *iters = nullptr,
                              StatusCode *status = nullptr)
{
    auto s = kin.screws();
    check(ppx_cuda_host_ik_solve(s.data(), kin.home().data(), (int)N, dp(init), dp(target), max_iter, dp(q_out), iters,
                                 reinterpret_cast<int8_t *>(status), n));
}

// linsolve<factorization::QR|SVD> over n systems: x[i] and (optionally) status[i]
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::QR>::type batch_linsolve(const Matrix<M, N> *A, const Matrix<M, 1> *b,
                                                                       Matrix<N, 1> *x, StatusCode *status,
                                                                       std::size_t n)
{
    check(ppx_cuda_host_linsolve_qr((int)M, (int)N, dp(A), dp(b), dp(x), reinterpret_cast<int8_t *>(status), n));
}
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::SVD>::type batch_linsolve(const Matrix<M, N> *A, const Matrix<M, 1> *b,
                                                                        Matrix<N, 1> *x, StatusCode *status,
                                                                        std::size_t n)
{
    check(ppx_cuda_host_linsolve_svd((int)M, (int)N, dp(A), dp(b), dp(x), n));
    if (status)
        std::fill(status, status + n, StatusCode::NORMAL);
}
template <Factorization type, std::size_t M, std::size_t N>
typename std::enable_if<type == Factorization::LU>::type batch_linsolve(const Matrix<M, N> *A, const Matrix<M, 1> *b,
                                                                       Matrix<N, 1> *x, StatusCode *status,
                                                                       std::size_t n)
{
    static_assert(M == N, "LU decomposation only supports square matrix.");
    check(ppx_cuda_host_linsolve_lu((int)N, dp(A), dp(b), dp(x), reinterpret_cast<int8_t *>(status), n));
}
template <std::size_t N>
void batch_inverse(const Matrix<N, N> *A, Matrix<N, N> *Ai, std::size_t n)
{
    check(ppx_cuda_host_inverse((int)N, dp(A), dp(Ai), n));
}
template <std::size_t N>
void batch_determinant(const Matrix<N, N> *A, double *det, std::size_t n)
{
    check(ppx_cuda_host_det((int)N, dp(A), det, n));
}
template <std::size_t M, std::size_t N>
void batch_pinv(const Matrix<M, N> *A, Matrix<N, M> *P, std::size_t n)
{
    check(ppx_cuda_host_pinv((int)M, (int)N, dp(A), dp(P), n));
}
template <std::size_t M, std::size_t N>
void batch_norm2(const Matrix<M, N> *A, double *s, std::size_t n)
{
    check(ppx_cuda_host_norm2((int)M, (int)N, dp(A), s, n));
}
// ... and into the reference's EqnResult<N> records
template <Factorization type, std::size_t M, std::size_t N>
void batch_linsolve(const Matrix<M, N> *A, const Matrix<M, 1> *b, EqnResult<N> *out, std::size_t n)
{
    std::vector<Matrix<N, 1>> x(n);
    std::vector<StatusCode> s(n);
    batch_linsolve<type, M, N>(A, b, x.data(), s.data(), n);
    for (std::size_t i = 0; i < n; i++)
    {
        out[i].x = x[i];
        out[i].s = s[i];
    }
}
// svdcmp over n matrices: U and/or V may be null
template <std::size_t M, std::size_t N>
void batch_svdcmp(const Matrix<M, N> *A, Matrix<M, N> *U, Matrix<N, 1> *w, Matrix<N, N> *V, std::size_t n)
{
    check(ppx_cuda_host_svd((int)M, (int)N, dp(A), dp(U), dp(w), dp(V), n));
}
// eig<eigensystem::...> over n symmetric matrices: vec may be null (SymOnlyVal)
template <eigensystem type, std::size_t N>
void batch_eig(const Matrix<N, N> *S, Matrix<N, 1> *val, Matrix<N, N> *vec, std::size_t n)
{
    check(ppx_cuda_host_eig_sym((int)N, dp(S), dp(val), type == eigensystem::SymOnlyVal ? nullptr : dp(vec),
                                type == eigensystem::SymValAndVec ? 0 : 1, n));
}
} // namespace cuda
} // namespace ppx
#endif
