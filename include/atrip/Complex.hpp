// Complex scalar type of the public API (reference src/atrip/Complex.hpp:16-52).  Atrip::run is
// instantiated for double and Complex, as in the reference (Atrip.cxx:1135-1136).
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
