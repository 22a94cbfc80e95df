// Public API of atrip, B200 build: atrip::Atrip::init / Input<F> / run<F> / Output.
//
// Same names, argument meaning and error behaviour as the reference's src/atrip/Atrip.hpp:41-140,
// so the reference's own driver (bench/main.cxx) compiles and runs against this header unchanged.
// run<double> drives the B200 engine through the C-ABI in include/atrip_b200.h; there is no CPU
// path.  Errors are thrown as std::string, like the reference's ACC layer (Acc.hpp:16-41).
#pragma once
#include <cstddef>
#include <map>
#include <string>

#include <mpi.h>

#include <atrip/Complex.hpp>
#include <atrip/Utils.hpp>

namespace atrip {

struct Atrip {
  // process-wide state, as in the reference (Atrip.hpp:43-51)
  static size_t rank;
  static size_t np;
  static MPI_Comm communicator;
  static std::map<std::string, double> chrono;  // seconds per phase of the last run

  static void init(MPI_Comm);

  template <typename F = double>
  struct Input {
    // non-owning tensor pointers (Atrip.hpp:69-72); Vppph is deleted by run when delete_Vppph
    CTF::Tensor<F> *ei = nullptr, *ea = nullptr, *Tph = nullptr, *Tpphh = nullptr, *Vpphh = nullptr,
                   *Vhhhp = nullptr, *Vppph = nullptr, *Jppph = nullptr, *Jhhhp = nullptr;
    Input &with_epsilon_i(CTF::Tensor<F> *t) { ei = t; return *this; }
    Input &with_epsilon_a(CTF::Tensor<F> *t) { ea = t; return *this; }
    Input &with_Tai(CTF::Tensor<F> *t) { Tph = t; return *this; }
    Input &with_Tabij(CTF::Tensor<F> *t) { Tpphh = t; return *this; }
    Input &with_Vabij(CTF::Tensor<F> *t) { Vpphh = t; return *this; }
    Input &with_Vijka(CTF::Tensor<F> *t) { Vhhhp = t; return *this; }
    Input &with_Vabci(CTF::Tensor<F> *t) { Vppph = t; return *this; }
    Input &with_Jijka(CTF::Tensor<F> *t) { Jhhhp = t; return *this; }
    Input &with_Jabci(CTF::Tensor<F> *t) { Jppph = t; return *this; }

    enum TuplesDistribution { NAIVE, GROUP_AND_SORT };

    // value attributes and their defaults (Atrip.hpp:113-131)
#define ATRIP_B200_ATTR(type, name, dflt) \
  type name = dflt;                       \
  Input &with_##name(type v) { name = v; return *this; }
    ATRIP_B200_ATTR(bool, delete_Vppph, false)
    ATRIP_B200_ATTR(bool, rank_round_robin, false)
    ATRIP_B200_ATTR(bool, chrono, false)
    ATRIP_B200_ATTR(bool, barrier, false)
    ATRIP_B200_ATTR(bool, blocking, false)
    ATRIP_B200_ATTR(size_t, max_iterations, 0)
    ATRIP_B200_ATTR(int, iteration_mod, -1)
    ATRIP_B200_ATTR(int, percentage_mod, -1)
    ATRIP_B200_ATTR(TuplesDistribution, tuples_distribution, NAIVE)
    ATRIP_B200_ATTR(std::string, checkpoint_path, "atrip-checkpoint.yaml")
    ATRIP_B200_ATTR(bool, read_checkpoint_if_exists, true)
    ATRIP_B200_ATTR(bool, writeCheckpoint, true)
    ATRIP_B200_ATTR(float, checkpoint_at_percentage, 10)
    ATRIP_B200_ATTR(size_t, checkpoint_at_every_iteration, 0)
    ATRIP_B200_ATTR(bool, ijkabc, 0)
    ATRIP_B200_ATTR(size_t, ooo_threads, 0)  // accepted and ignored, as in the reference (B7)
    ATRIP_B200_ATTR(size_t, ooo_blocks, 0)
#undef ATRIP_B200_ATTR
  };

  struct Output {
    double energy;
    double ct_energy;
  };

  template <typename F = double>
  static Output run(Input<F> const &in);
};

}  // namespace atrip
