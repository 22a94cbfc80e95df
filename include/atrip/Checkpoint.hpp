// Checkpoint file of a (T) run: same "Key: value" text format as the reference
// (src/atrip/Checkpoint.hpp:27-79), so files are interchangeable.  One optional extra key,
// "CtEnergy", carries the (cT) partial sum (the reference's reader skips keys it does not know; its
// checkpoint simply loses ct_energy on a resume).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <string>

namespace atrip {

struct Checkpoint {
  size_t no = 0, nv = 0, nranks = 0, nnodes = 0;
  double energy = 0;
  size_t iteration = 0;
  bool rank_round_robin = false;
  double ct_energy = 0;  // only written / valid when has_ct
  bool has_ct = false;
};

inline void write_checkpoint(Checkpoint const &c, std::string const &path) {
  std::ofstream f(path);
  f << "No: " << c.no << "\nNv: " << c.nv << "\nNranks: " << c.nranks << "\nNnodes: " << c.nnodes
    << "\nEnergy: " << std::setprecision(19) << c.energy << "\nIteration: " << c.iteration
    << "\nRankRoundRobin: " << (c.rank_round_robin ? "true" : "false") << "\n";
  if (c.has_ct) f << "CtEnergy: " << std::setprecision(19) << c.ct_energy << "\n";
}

inline Checkpoint read_checkpoint(std::ifstream &f) {
  Checkpoint c;
  std::string line;
  while (std::getline(f, line)) {
    const auto colon = line.find(':');
    if (colon == std::string::npos) continue;
    auto strip = [](std::string s) {
      const auto b = s.find_first_not_of(" \t"), e = s.find_last_not_of(" \t\r");
      return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    };
    const std::string key = strip(line.substr(0, colon)), val = strip(line.substr(colon + 1));
    if (key == "No") c.no = std::strtoull(val.c_str(), nullptr, 10);
    else if (key == "Nv") c.nv = std::strtoull(val.c_str(), nullptr, 10);
    else if (key == "Nranks") c.nranks = std::strtoull(val.c_str(), nullptr, 10);
    else if (key == "Nnodes") c.nnodes = std::strtoull(val.c_str(), nullptr, 10);
    else if (key == "Energy") c.energy = std::strtod(val.c_str(), nullptr);
    else if (key == "Iteration") c.iteration = std::strtoull(val.c_str(), nullptr, 10);
    else if (key == "RankRoundRobin") c.rank_round_robin = !val.empty() && val[0] == 't';
    else if (key == "CtEnergy") {
      c.ct_energy = std::strtod(val.c_str(), nullptr);
      c.has_ct = true;
    }
  }
  return c;
}

inline Checkpoint read_checkpoint(std::string const &path) {
  std::ifstream f(path);
  return read_checkpoint(f);
}

}  // namespace atrip
