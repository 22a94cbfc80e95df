// Scalar helpers of the public API that the reference's driver uses
// (reference src/atrip/Operations.hpp:28-53; bench/main.cxx:326-328).  Host-only here: the device
// code applies the conjugations of the complex field itself (stores.cuh: -conj(Vijka) in the AX
// slices; reduction_z.cuh: conj(Tijk) in the energy sums).
#pragma once
#include <atrip/Complex.hpp>

namespace atrip {
namespace acc {
template <typename F> inline F maybe_conjugate_scalar(F const &x) { return x; }
template <> inline Complex maybe_conjugate_scalar(Complex const &x) { return std::conj(x); }
template <typename F> inline F prod(F const &a, F const &b) { return a * b; }
template <typename F> inline F div(F const &a, F const &b) { return a / b; }
template <typename F> inline F add(F const &a, F const &b) { return a + b; }
template <typename F> inline F sub(F const &a, F const &b) { return a - b; }
template <typename F> inline double real(F const &a) { return std::real(a); }
template <typename F> inline void sum_in_place(F *to, F const *from) { *to += *from; }
}  // namespace acc
}  // namespace atrip
