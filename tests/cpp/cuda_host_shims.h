// cuda_host_shims.h -- TEST HARNESS support: the handful of CUDA spellings the device headers under
// matrix_b200/csrc use outside templates, defined for a plain host compiler so that the per-lane device math can be
// compiled with g++ and replayed in the CPU test-suite (tests/cpp/codo_host.cpp, tests/cpp/math_host.cpp).  A "warp"
// is one lane here: the votes return their own predicate.  Not part of the product.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
static inline int __double2hiint(double x)
{
    int64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(u >> 32);
}
static inline int __double2loint(double x)
{
    int64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(u & 0xffffffff);
}
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    std::memcpy(&x, &u, 8);
    return x;
}
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
