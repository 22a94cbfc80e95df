// Dense, single-process stand-in for Cyclops (<ctf.hpp>).
//
// atrip's whole CTF surface is data movement: Tensor(order, lens, syms, World),
// lens, data, read_all, slice  (reference Atrip.cxx:72-73,179-181,
// Unions.hpp:49-65, SliceUnion.cxx:308-317) plus, in bench/main.cxx, World(argc,
// argv), fill_random, read_dense_from_file, index expressions, Transform and
// norm2 (bench/main.cxx:44-89, 224-233, 300-312).  This shim provides exactly
// that on one process with column-major (first index fastest) dense storage, so
// that the host mirror in this repo, the reference's own bench driver and the
// oracle build of the reference sources compile in an image without CTF.
// With a real CTF on the include path this file is not used.
//
// Like the real ctf.hpp it leaks `using namespace std`: the reference's
// Tuples.hpp:68 relies on that for an unqualified `string`.
#ifndef ATRIP_B200_DENSE_CTF_HPP
#define ATRIP_B200_DENSE_CTF_HPP

#include <mpi.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

using namespace std;

enum { NS = 0, SY = 1, AS = 2, SH = 3 };

namespace CTF {

struct World {
  MPI_Comm comm;
  int rank, np;
  World(MPI_Comm c = MPI_COMM_WORLD) : comm(c) {
    MPI_Comm_rank(c, &rank);
    MPI_Comm_size(c, &np);
  }
  World(int, char **) : World(MPI_COMM_WORLD) {}
};

namespace detail {
inline double real_part(double x) { return x; }
inline double real_part(std::complex<double> const &x) { return x.real(); }
inline double abs2(double x) { return x * x; }
inline double abs2(std::complex<double> const &x) { return std::norm(x); }
// counter-free xorshift for fill_random (NOT CTF's generator: synthetic bench
// inputs are timing-only; parity inputs never come from here)
inline uint64_t next(uint64_t &s) {
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return s;
}
} // namespace detail

template <typename F>
class Tensor;

template <typename F>
struct Idx_Tensor {
  Tensor<F> *t;
  std::string idx;
  void operator=(Idx_Tensor<F> const &rhs);
};

template <typename F = double>
class Tensor {
public:
  int order = 0;
  int64_t *lens = nullptr;
  int *sym = nullptr;
  World *wrld = nullptr;
  F *data = nullptr;
  int64_t size = 0;

  Tensor() {}
  Tensor(int order_, int const *lens_, int const *sym_, World &w)
      : order(order_), wrld(&w) {
    lens = new int64_t[order > 0 ? order : 1];
    sym = new int[order > 0 ? order : 1];
    size = 1;
    for (int i = 0; i < order; i++) {
      lens[i] = lens_[i];
      sym[i] = sym_ ? sym_[i] : NS;
      size *= lens[i];
    }
    data = new F[size]();
  }
  Tensor(Tensor const &o) : order(o.order), wrld(o.wrld), size(o.size) {
    lens = new int64_t[order > 0 ? order : 1];
    sym = new int[order > 0 ? order : 1];
    for (int i = 0; i < order; i++) {
      lens[i] = o.lens[i];
      sym[i] = o.sym[i];
    }
    data = new F[size];
    std::copy(o.data, o.data + size, data);
  }
  Tensor &operator=(Tensor const &) = delete;
  ~Tensor() {
    delete[] lens;
    delete[] sym;
    delete[] data;
  }

  void read_all(F *out) const { std::copy(data, data + size, out); }

  void fill_random(F a, F b) {
    uint64_t s = 0x9E3779B97F4A7C15ull ^ (uint64_t)(uintptr_t)this ^ (uint64_t)size;
    const double lo = detail::real_part(a), hi = detail::real_part(b);
    for (int64_t i = 0; i < size; i++) {
      const double u = (double)(detail::next(s) >> 11) * (1.0 / 9007199254740992.0);
      data[i] = F(lo + (hi - lo) * u);
    }
  }

  // raw native-endian F array in global column-major order
  void read_dense_from_file(char const *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) throw std::string("ctf shim: cannot open ") + path;
    f.read(reinterpret_cast<char *>(data), sizeof(F) * size);
    if (f.gcount() != (std::streamsize)(sizeof(F) * size))
      throw std::string("ctf shim: short read on ") + path;
  }

  double norm2() const {
    double s = 0;
    for (int64_t i = 0; i < size; i++) s += detail::abs2(data[i]);
    return std::sqrt(s);
  }

  // this[low,up) = beta * this[low,up) + alpha * A[lowA,upA)
  // The two boxes may have different orders (singleton dimensions dropped);
  // elements correspond by the column-major enumeration of each box.
  void slice(int const *low, int const *up, F beta, Tensor const &A,
             int const *lowA, int const *upA, F alpha) {
    std::vector<int64_t> ext(order > 0 ? order : 1, 1), extA(A.order > 0 ? A.order : 1, 1);
    int64_t n = 1, nA = 1;
    for (int d = 0; d < order; d++) n *= (ext[d] = up[d] - low[d]);
    for (int d = 0; d < A.order; d++) nA *= (extA[d] = upA[d] - lowA[d]);
    if (n != nA) throw std::string("ctf shim: slice boxes differ in size");
    std::vector<int64_t> c(order > 0 ? order : 1, 0), cA(A.order > 0 ? A.order : 1, 0);
    for (int64_t e = 0; e < n; e++) {
      int64_t off = 0, offA = 0, stride = 1;
      for (int d = 0; d < order; d++) {
        off += (low[d] + c[d]) * stride;
        stride *= lens[d];
      }
      stride = 1;
      for (int d = 0; d < A.order; d++) {
        offA += (lowA[d] + cA[d]) * stride;
        stride *= A.lens[d];
      }
      data[off] = beta * data[off] + alpha * A.data[offA];
      for (int d = 0; d < order; d++) {
        if (++c[d] < ext[d]) break;
        c[d] = 0;
      }
      for (int d = 0; d < A.order; d++) {
        if (++cA[d] < extA[d]) break;
        cA[d] = 0;
      }
    }
  }

  Idx_Tensor<F> operator[](char const *idx) { return Idx_Tensor<F>{this, idx}; }
};

namespace detail {
// call f(offset_in_a, offset_in_b) for every joint index assignment, where the
// index strings name the modes of a and b (all labels of b must occur in a)
template <typename FA, typename FB, typename Fn>
void for_each_matched(Tensor<FA> &a, std::string const &ia, Tensor<FB> &b,
                      std::string const &ib, Fn f) {
  const int n = a.order;
  std::vector<int64_t> c(n > 0 ? n : 1, 0), sb(n > 0 ? n : 1, 0);
  for (int d = 0; d < n; d++) {
    const size_t p = ib.find(ia[d]);
    if (p == std::string::npos) throw std::string("ctf shim: unmatched index");
    int64_t s = 1;
    for (size_t q = 0; q < p; q++) s *= b.lens[q];
    sb[d] = s;
  }
  for (int64_t e = 0; e < a.size; e++) {
    int64_t ob = 0;
    for (int d = 0; d < n; d++) ob += c[d] * sb[d];
    f(e, ob);
    for (int d = 0; d < n; d++) {
      if (++c[d] < a.lens[d]) break;
      c[d] = 0;
    }
  }
}
} // namespace detail

template <typename F>
void Idx_Tensor<F>::operator=(Idx_Tensor<F> const &rhs) {
  detail::for_each_matched(*t, idx, *rhs.t, rhs.idx,
                           [&](int64_t o, int64_t r) { t->data[o] = rhs.t->data[r]; });
}

template <typename A, typename B = A>
struct Transform {
  std::function<void(A, B &)> fn;
  template <typename L>
  Transform(L l) : fn(l) {}
  void operator()(Idx_Tensor<A> a, Idx_Tensor<B> b) const {
    detail::for_each_matched(*b.t, b.idx, *a.t, a.idx, [&](int64_t ob, int64_t oa) {
      fn(a.t->data[oa], b.t->data[ob]);
    });
  }
};

} // namespace CTF

#endif
