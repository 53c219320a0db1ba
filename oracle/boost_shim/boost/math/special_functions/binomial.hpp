// Shim for boost::math::binomial_coefficient<double>(n, k), the only Boost.Math symbol the
// reference uses (cpp/src/hermite/iterators.cpp:265-284).  Boost rounds its floating-point result to
// the nearest integer (ceil(x - 0.5)); an exact integer evaluation therefore returns the same value
// wherever Boost's own result is within 0.5 of the true binomial (all values below 2^53).
#pragma once
namespace boost { namespace math {
template <typename T>
inline T binomial_coefficient(unsigned n, unsigned k) {
    if (k > n) return T(0);
    if (k > n - k) k = n - k;
    long double r = 1.0L;
    for (unsigned i = 1; i <= k; i++) {
        r = r * (long double)(n - k + i) / (long double)i;  // exact: product of i consecutive ints / i!
        r = (long double)(unsigned long long)(r + 0.5L);
    }
    return (T)r;
}
}}
