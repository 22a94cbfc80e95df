// Example/parity driver of the C++ API: builds CTF tensors holding the counter-based synthetic
// inputs, calls atrip::Atrip::run<double> exactly like the reference's bench
// (bench/main.cxx:345-391) and prints the energies in hex and decimal.
//   synth_driver <No> <Nv> <seed> <scale> [max_iterations] [dist: group|naive] [cT|T] [complex|real] [ijkabc]
// With "complex" the tensors are CTF::Tensor<Complex> and atrip::Atrip::run<Complex> is called:
// element e of a complex tensor is (synth(2e), synth(2e+1)), the epsilons are (synth(e), 0).
// "ijkabc" sets Input::ijkabc (Atrip.cxx:183-187, 1108-1111).
// Several ranks, one per GPU: start one process per rank with RANK / WORLD_SIZE / ATRIP_SHIM_MPI=1 and a
// fresh ATRIP_SHIM_MPI_DIR in the environment (include/shim/mpi.h; tools/launch_ranks.py does it), or
// build against a real MPI and use mpirun.  Rank 0 prints the RESULT line.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <atrip.hpp>
#include <atrip_b200.h>

static CTF::Tensor<double> *make(CTF::World &w, std::vector<int> lens, int tensor_id, uint64_t seed, double scale) {
  std::vector<int> syms(lens.size(), NS);
  auto *t = new CTF::Tensor<double>((int)lens.size(), lens.data(), syms.data(), w);
  if (atrip_b200_synth_to_host(w.rank % std::max(1, atrip_b200_device_count()), seed, tensor_id, scale, 0, (uint64_t)t->size, t->data) != 0) {
    std::fprintf(stderr, "synth failed: %s\n", atrip_b200_last_error());
    std::exit(2);
  }
  return t;
}

static CTF::Tensor<atrip::Complex> *make_z(CTF::World &w, std::vector<int> lens, int tensor_id, uint64_t seed,
                                           double scale) {
  std::vector<int> syms(lens.size(), NS);
  auto *t = new CTF::Tensor<atrip::Complex>((int)lens.size(), lens.data(), syms.data(), w);
  double *d = reinterpret_cast<double *>(t->data);
  const uint64_t n = (uint64_t)t->size;
  const bool eps = tensor_id <= 1;
  if (atrip_b200_synth_to_host(w.rank % std::max(1, atrip_b200_device_count()), seed, tensor_id, scale, 0, eps ? n : 2 * n, d) != 0) {
    std::fprintf(stderr, "synth failed: %s\n", atrip_b200_last_error());
    std::exit(2);
  }
  if (eps)  // n real values were written to the front: spread them to (value, 0) pairs
    for (uint64_t e = n; e-- > 0;) {
      d[2 * e] = d[e];
      d[2 * e + 1] = 0.0;
    }
  return t;
}

template <typename F, typename Make>
static int go(CTF::World &world, Make mk, int No, int Nv, uint64_t seed, double scale, size_t max_it, bool naive,
              bool cT, bool ijkabc) {
  using In = atrip::Atrip::Input<F>;
  auto in = In()
                .with_epsilon_i(mk(world, {No}, 0, seed, scale))
                .with_epsilon_a(mk(world, {Nv}, 1, seed, scale))
                .with_Tai(mk(world, {Nv, No}, 2, seed, scale))
                .with_Tabij(mk(world, {Nv, Nv, No, No}, 3, seed, scale))
                .with_Vabij(mk(world, {Nv, Nv, No, No}, 4, seed, scale))
                .with_Vijka(mk(world, {No, No, No, Nv}, 5, seed, scale))
                .with_Vabci(mk(world, {Nv, Nv, Nv, No}, 6, seed, scale))
                .with_Jijka(cT ? mk(world, {No, No, No, Nv}, 7, seed, scale) : nullptr)
                .with_Jabci(cT ? mk(world, {Nv, Nv, Nv, No}, 8, seed, scale) : nullptr)
                .with_delete_Vppph(true)
                .with_tuples_distribution(naive ? In::NAIVE : In::GROUP_AND_SORT)
                .with_max_iterations(max_it)
                .with_ijkabc(ijkabc)
                .with_read_checkpoint_if_exists(false)
                .with_writeCheckpoint(false)
                .with_percentage_mod(25);
  // SYNTH_CHECKPOINT=<path> [SYNTH_CHECKPOINT_EVERY=<iterations>]: write checkpoints and resume from <path>
  // if it exists (Atrip.cxx:586-621, 713-731)
  if (const char *ck = std::getenv("SYNTH_CHECKPOINT")) {
    const char *ev = std::getenv("SYNTH_CHECKPOINT_EVERY");
    in.with_checkpoint_path(ck)
        .with_read_checkpoint_if_exists(true)
        .with_writeCheckpoint(true)
        .with_checkpoint_at_every_iteration(ev ? std::strtoull(ev, nullptr, 10) : 10);
  }
  try {
    auto out = atrip::Atrip::run<F>(in);
    if (world.rank == 0)
      std::printf("RESULT energy %a %.17g ct_energy %a %.17g\n", out.energy, out.energy, out.ct_energy, out.ct_energy);
  } catch (std::string const &m) {
    std::printf("Atrip throwed with msg: %s\n", m.c_str());
    return 1;
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s No Nv seed scale [max_iterations] [group|naive] [cT|T] [complex|real] [ijkabc]\n", argv[0]);
    return 2;
  }
  const int No = std::atoi(argv[1]), Nv = std::atoi(argv[2]);
  const uint64_t seed = std::strtoull(argv[3], nullptr, 10);
  const double scale = std::atof(argv[4]);
  const size_t max_it = argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 0;
  const bool naive = argc > 6 && !std::strcmp(argv[6], "naive");
  const bool cT = argc > 7 && !std::strcmp(argv[7], "cT");
  MPI_Init(&argc, &argv);
  CTF::World world(argc, argv);
  atrip::Atrip::init(world.comm);
  const bool cplx = argc > 8 && !std::strcmp(argv[8], "complex");
  const bool ijkabc = argc > 9 && !std::strcmp(argv[9], "ijkabc");
  const int rc = cplx ? go<atrip::Complex>(world, make_z, No, Nv, seed, scale, max_it, naive, cT, ijkabc)
                      : go<double>(world, make, No, Nv, seed, scale, max_it, naive, cT, ijkabc);
  if (rc) return rc;
  MPI_Finalize();
  return 0;
}
