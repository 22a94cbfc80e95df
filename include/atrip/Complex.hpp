// Complex scalar type of the public API (reference src/atrip/Complex.hpp:16-52).  The B200 engine
// computes the FP64 real case only (BASELINE.json north_star); run<Complex> throws.
#pragma once
#include <complex>
#include <type_traits>

namespace atrip {
using Complex = std::complex<double>;
namespace traits {
template <typename F> struct is_complex : std::false_type {};
template <> struct is_complex<Complex> : std::true_type {};
}  // namespace traits
}  // namespace atrip
