// Example/parity driver of the C++ API: builds CTF tensors holding the counter-based synthetic
// inputs, calls atrip::Atrip::run<double> exactly like the reference's bench
// (bench/main.cxx:345-391) and prints the energies in hex and decimal.
//   synth_driver <No> <Nv> <seed> <scale> [max_iterations] [dist: group|naive] [cT]
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <atrip.hpp>
#include <atrip_b200.h>

static CTF::Tensor<double> *make(CTF::World &w, std::vector<int> lens, int tensor_id, uint64_t seed, double scale) {
  std::vector<int> syms(lens.size(), NS);
  auto *t = new CTF::Tensor<double>((int)lens.size(), lens.data(), syms.data(), w);
  if (atrip_b200_synth_to_host(0, seed, tensor_id, scale, 0, (uint64_t)t->size, t->data) != 0) {
    std::fprintf(stderr, "synth failed: %s\n", atrip_b200_last_error());
    std::exit(2);
  }
  return t;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s No Nv seed scale [max_iterations] [group|naive] [cT]\n", argv[0]);
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
  using In = atrip::Atrip::Input<double>;
  auto in = In()
                .with_epsilon_i(make(world, {No}, 0, seed, scale))
                .with_epsilon_a(make(world, {Nv}, 1, seed, scale))
                .with_Tai(make(world, {Nv, No}, 2, seed, scale))
                .with_Tabij(make(world, {Nv, Nv, No, No}, 3, seed, scale))
                .with_Vabij(make(world, {Nv, Nv, No, No}, 4, seed, scale))
                .with_Vijka(make(world, {No, No, No, Nv}, 5, seed, scale))
                .with_Vabci(make(world, {Nv, Nv, Nv, No}, 6, seed, scale))
                .with_Jijka(cT ? make(world, {No, No, No, Nv}, 7, seed, scale) : nullptr)
                .with_Jabci(cT ? make(world, {Nv, Nv, Nv, No}, 8, seed, scale) : nullptr)
                .with_delete_Vppph(true)
                .with_tuples_distribution(naive ? In::NAIVE : In::GROUP_AND_SORT)
                .with_max_iterations(max_it)
                .with_read_checkpoint_if_exists(false)
                .with_writeCheckpoint(false)
                .with_percentage_mod(25);
  try {
    auto out = atrip::Atrip::run<double>(in);
    std::printf("RESULT energy %a %.17g ct_energy %a %.17g\n", out.energy, out.energy, out.ct_energy, out.ct_energy);
  } catch (std::string const &m) {
    std::printf("Atrip throwed with msg: %s\n", m.c_str());
    return 1;
  }
  MPI_Finalize();
  return 0;
}
