// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" entry points over the reference's OWN, UNMODIFIED implementation
// (compiled from /root/reference/src/atrip/*.cxx where it lies, see
// oracle/Makefile) so that tests and bench.py's cpu_baseline can call it
// through ctypes.  Nothing here restates the algorithm: every function just
// forwards to the reference symbol named in its comment.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <atrip.hpp>
#include <atrip/Equations.hpp>
#include <atrip/Tuples.hpp>

using namespace atrip;

namespace {
template <typename F = double>
CTF::Tensor<F> *wrap(CTF::World &w, std::vector<int> lens, const F *src) {
  std::vector<int> syms(lens.size(), NS);
  auto *t = new CTF::Tensor<F>((int)lens.size(), lens.data(), syms.data(), w);
  std::memcpy(t->data, src, sizeof(F) * t->size);
  return t;
}
bool initialised = false;
} // namespace

// Input::ijkabc of the following ref_run / ref_run_z calls (reference Atrip.cxx:183-187, 1108-1111:
// Tai is negated and the final sign flip is skipped)
static bool g_ijkabc = false;

extern "C" {

void ref_set_ijkabc(int on) { g_ijkabc = on != 0; }

// atrip::Atrip::init + atrip::Atrip::run<double> (reference Atrip.cxx:54-63,
// 65-1133) at np = 1 with GROUP_AND_SORT (the NAIVE distribution is broken at
// reference HEAD).  Tensors are column-major with the
// lens of reference bench/main.cxx:198-200.  Returns 0, or 1 if the reference
// threw (message copied to err).
int ref_run(int No, int Nv, const double *epsi, const double *epsa,
            const double *Tai, const double *Tabij, const double *Vabij,
            const double *Vijka, const double *Vabci, const double *Jijka,
            const double *Jabci, long max_iterations, double *energy,
            double *ct_energy, char *err, int errlen) {
  try {
    CTF::World world(MPI_COMM_WORLD);
    if (!initialised) {
      Atrip::init(world.comm);
      initialised = true;
    }
    Atrip::chrono.clear();
    auto *ei = wrap(world, {No}, epsi);
    auto *ea = wrap(world, {Nv}, epsa);
    auto *tph = wrap(world, {Nv, No}, Tai);
    auto *tpphh = wrap(world, {Nv, Nv, No, No}, Tabij);
    auto *vpphh = wrap(world, {Nv, Nv, No, No}, Vabij);
    auto *vhhhp = wrap(world, {No, No, No, Nv}, Vijka);
    auto *vppph = wrap(world, {Nv, Nv, Nv, No}, Vabci);
    CTF::Tensor<double> *jhhhp = Jijka ? wrap(world, {No, No, No, Nv}, Jijka) : nullptr;
    CTF::Tensor<double> *jppph = Jabci ? wrap(world, {Nv, Nv, Nv, No}, Jabci) : nullptr;
    auto in = Atrip::Input<double>()
                  .with_epsilon_i(ei)
                  .with_epsilon_a(ea)
                  .with_Tai(tph)
                  .with_Tabij(tpphh)
                  .with_Vabij(vpphh)
                  .with_Vijka(vhhhp)
                  .with_Vabci(vppph)
                  .with_Jijka(jhhhp)
                  .with_Jabci(jppph)
                  .with_delete_Vppph(false)
                  .with_ijkabc(g_ijkabc)
                  .with_tuples_distribution(
                      Atrip::Input<double>::TuplesDistribution::GROUP_AND_SORT)
                  .with_max_iterations((size_t)max_iterations)
                  .with_iteration_mod(-1)
                  .with_percentage_mod(-1)
                  .with_read_checkpoint_if_exists(false)
                  // keeps checkpoint_mod != 0
                  .with_checkpoint_at_every_iteration((size_t)1 << 60);
    auto out = Atrip::run<double>(in);
    *energy = out.energy;
    *ct_energy = out.ct_energy;
    delete ei;
    delete ea;
    delete tph;
    delete tpphh;
    delete vpphh;
    delete vhhhp;
    delete vppph;
    delete jhhhp;
    delete jppph;
    return 0;
  } catch (const char *m) {
    std::strncpy(err, m, errlen - 1);
  } catch (std::string const &m) {
    std::strncpy(err, m.c_str(), errlen - 1);
  } catch (std::exception const &e) {
    std::strncpy(err, e.what(), errlen - 1);
  }
  return 1;
}

// seconds spent in the reference's own timers during the last ref_run
// (Atrip::chrono, reference Chrono.hpp); 0 if the timer was never started
double ref_chrono(const char *name) {
  auto it = Atrip::chrono.find(name);
  return it == Atrip::chrono.end() ? 0.0 : it->second.count();
}

// atrip::doubles_contribution<double> (reference Equations.cxx:455-728); the
// two No^3 scratch buffers the reference wants are provided by the caller
void ref_doubles(long No, long Nv, double *VAB, double *VAC, double *VBC,
                 double *VBA, double *VCA, double *VCB, double *HA, double *HB,
                 double *HC, double *TA, double *TB, double *TC, double *TAB,
                 double *TAC, double *TBC, double *Tijk, double *tbuf,
                 double *vhhh) {
  doubles_contribution<double>((size_t)No, (size_t)Nv, VAB, VAC, VBC, VBA, VCA,
                               VCB, HA, HB, HC, TA, TB, TC, TAB, TAC, TBC, Tijk,
                               tbuf, vhhh);
}

// atrip::singles_contribution<double> (reference Equations.cxx:387-426)
void ref_singles(long No, long Nv, long a, long b, long c, double *Tph,
                 double *VABij, double *VACij, double *VBCij, double *Zijk) {
  singles_contribution<double>((size_t)No, (size_t)Nv, (size_t)a, (size_t)b,
                               (size_t)c, Tph, VABij, VACij, VBCij, Zijk);
}

// atrip::get_energy_distinct / get_energy_same (reference Equations.cxx:101-238)
double ref_energy_distinct(double epsabc, long No, double *epsi, double *Tijk,
                           double *Zijk) {
  double e = 0;
  get_energy_distinct<double>(epsabc, (size_t)No, epsi, Tijk, Zijk, &e);
  return e;
}
double ref_energy_same(double epsabc, long No, double *epsi, double *Tijk,
                       double *Zijk) {
  double e = 0;
  get_energy_same<double>(epsabc, (size_t)No, epsi, Tijk, Zijk, &e);
  return e;
}

// ---- F = Complex (reference instantiations Atrip.cxx:1136, Equations.cxx:730-795).
// Arrays are interleaved (re, im) doubles = std::complex<double> memory layout.

// atrip::Atrip::run<Complex> (reference Atrip.cxx:65-1133), same options as ref_run
int ref_run_z(int No, int Nv, const double *epsi, const double *epsa,
              const double *Tai, const double *Tabij, const double *Vabij,
              const double *Vijka, const double *Vabci, const double *Jijka,
              const double *Jabci, long max_iterations, double *energy,
              double *ct_energy, char *err, int errlen) {
  try {
    CTF::World world(MPI_COMM_WORLD);
    if (!initialised) {
      Atrip::init(world.comm);
      initialised = true;
    }
    Atrip::chrono.clear();
    auto Z = [](const double *p) { return reinterpret_cast<const Complex *>(p); };
    auto *ei = wrap<Complex>(world, {No}, Z(epsi));
    auto *ea = wrap<Complex>(world, {Nv}, Z(epsa));
    auto *tph = wrap<Complex>(world, {Nv, No}, Z(Tai));
    auto *tpphh = wrap<Complex>(world, {Nv, Nv, No, No}, Z(Tabij));
    auto *vpphh = wrap<Complex>(world, {Nv, Nv, No, No}, Z(Vabij));
    auto *vhhhp = wrap<Complex>(world, {No, No, No, Nv}, Z(Vijka));
    auto *vppph = wrap<Complex>(world, {Nv, Nv, Nv, No}, Z(Vabci));
    CTF::Tensor<Complex> *jhhhp = Jijka ? wrap<Complex>(world, {No, No, No, Nv}, Z(Jijka)) : nullptr;
    CTF::Tensor<Complex> *jppph = Jabci ? wrap<Complex>(world, {Nv, Nv, Nv, No}, Z(Jabci)) : nullptr;
    auto in = Atrip::Input<Complex>()
                  .with_epsilon_i(ei)
                  .with_epsilon_a(ea)
                  .with_Tai(tph)
                  .with_Tabij(tpphh)
                  .with_Vabij(vpphh)
                  .with_Vijka(vhhhp)
                  .with_Vabci(vppph)
                  .with_Jijka(jhhhp)
                  .with_Jabci(jppph)
                  .with_delete_Vppph(false)
                  .with_ijkabc(g_ijkabc)
                  .with_tuples_distribution(
                      Atrip::Input<Complex>::TuplesDistribution::GROUP_AND_SORT)
                  .with_max_iterations((size_t)max_iterations)
                  .with_iteration_mod(-1)
                  .with_percentage_mod(-1)
                  .with_read_checkpoint_if_exists(false)
                  .with_checkpoint_at_every_iteration((size_t)1 << 60);
    auto out = Atrip::run<Complex>(in);
    *energy = out.energy;
    *ct_energy = out.ct_energy;
    delete ei;
    delete ea;
    delete tph;
    delete tpphh;
    delete vpphh;
    delete vhhhp;
    delete vppph;
    delete jhhhp;
    delete jppph;
    return 0;
  } catch (const char *m) {
    std::strncpy(err, m, errlen - 1);
  } catch (std::string const &m) {
    std::strncpy(err, m.c_str(), errlen - 1);
  } catch (std::exception const &e) {
    std::strncpy(err, e.what(), errlen - 1);
  }
  return 1;
}

// atrip::doubles_contribution<Complex> (reference Equations.cxx:455-728)
void ref_doubles_z(long No, long Nv, double *VAB, double *VAC, double *VBC,
                   double *VBA, double *VCA, double *VCB, double *HA, double *HB,
                   double *HC, double *TA, double *TB, double *TC, double *TAB,
                   double *TAC, double *TBC, double *Tijk, double *tbuf,
                   double *vhhh) {
  auto Z = [](double *p) { return reinterpret_cast<Complex *>(p); };
  doubles_contribution<Complex>((size_t)No, (size_t)Nv, Z(VAB), Z(VAC), Z(VBC), Z(VBA), Z(VCA),
                                Z(VCB), Z(HA), Z(HB), Z(HC), Z(TA), Z(TB), Z(TC), Z(TAB), Z(TAC),
                                Z(TBC), Z(Tijk), Z(tbuf), Z(vhhh));
}

// atrip::singles_contribution<Complex> (reference Equations.cxx:387-426)
void ref_singles_z(long No, long Nv, long a, long b, long c, double *Tph,
                   double *VABij, double *VACij, double *VBCij, double *Zijk) {
  auto Z = [](double *p) { return reinterpret_cast<Complex *>(p); };
  singles_contribution<Complex>((size_t)No, (size_t)Nv, (size_t)a, (size_t)b, (size_t)c, Z(Tph),
                                Z(VABij), Z(VACij), Z(VBCij), Z(Zijk));
}

// atrip::get_energy_distinct / get_energy_same <Complex> (reference Equations.cxx:101-238)
double ref_energy_distinct_z(double epsabc, long No, double *epsi, double *Tijk, double *Zijk) {
  auto Z = [](double *p) { return reinterpret_cast<Complex *>(p); };
  double e = 0;
  get_energy_distinct<Complex>(Complex(epsabc), (size_t)No, Z(epsi), Z(Tijk), Z(Zijk), &e);
  return e;
}
double ref_energy_same_z(double epsabc, long No, double *epsi, double *Tijk, double *Zijk) {
  auto Z = [](double *p) { return reinterpret_cast<Complex *>(p); };
  double e = 0;
  get_energy_same<Complex>(Complex(epsabc), (size_t)No, Z(epsi), Z(Tijk), Z(Zijk), &e);
  return e;
}

// atrip::group_and_sort::special_distribution over get_all_tuples_list(Nv)
// (reference Tuples.cxx:122-134, 156-308).  Writes up to cap tuples (3 x
// uint64 each) and returns the node's tuple count.
long ref_group_and_sort(long n_nodes, long node_id, long Nv, uint64_t *out, long cap) {
  auto const all = get_all_tuples_list((size_t)Nv);
  auto const mine = group_and_sort::special_distribution(
      group_and_sort::Info{(size_t)n_nodes, (size_t)node_id}, all);
  long n = (long)mine.size();
  for (long i = 0; i < n && i < cap; i++)
    for (int d = 0; d < 3; d++) out[3 * i + d] = mine[i][d];
  return n;
}

// atrip::get_all_tuples_list (reference Tuples.cxx:122-134)
long ref_all_tuples(long Nv, uint64_t *out, long cap) {
  auto const all = get_all_tuples_list((size_t)Nv);
  long n = (long)all.size();
  for (long i = 0; i < n && i < cap; i++)
    for (int d = 0; d < 3; d++) out[3 * i + d] = all[i][d];
  return n;
}

#if !defined(ATRIP_USE_DGEMM)
// The loop build (config without ATRIP_USE_DGEMM, reference Equations.cxx:685-727)
// never calls dgemm_, but Blas.cxx still references the BLAS symbols; satisfy
// the linker without a BLAS.  dcopy_ IS used (Atrip.cxx:899-906, Zijk = Tijk).
void dgemm_(const char *, const char *, const int *, const int *, const int *,
            double *, const double *, const int *, const double *, const int *,
            double *, double *, const int *) {
  std::abort();
}
void zgemm_(const char *, const char *, const int *, const int *, const int *,
            Complex *, const Complex *, const int *, const Complex *, const int *,
            Complex *, Complex *, const int *) {
  std::abort();
}
void dcopy_(int *n, const double *x, int *incx, double *y, int *incy) {
  for (int i = 0; i < *n; i++) y[(long)i * *incy] = x[(long)i * *incx];
}
void zcopy_(int *n, const void *x, int *incx, void *y, int *incy) {
  auto *xs = (const double *)x;
  auto *ys = (double *)y;
  for (int i = 0; i < *n; i++) {
    ys[2L * i * *incy] = xs[2L * i * *incx];
    ys[2L * i * *incy + 1] = xs[2L * i * *incx + 1];
  }
}
#endif

} // extern "C"
