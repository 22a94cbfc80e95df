// Logging and the per-report user callback of the public API
// (reference src/atrip/Debug.hpp:81-109; used by bench/main.cxx:124-142).
#pragma once
#include <cstddef>
#include <functional>
#include <iostream>

#ifndef LOG
#  ifdef ATRIP_NO_OUTPUT
#    define LOG(level, name) if (false) std::cout << name << ": "
#  else
#    define LOG(level, name) if (atrip::Atrip::rank == 0) std::cout << name << ": "
#  endif
#endif

namespace atrip {

struct IterationDescription;
using IterationDescriptor = std::function<void(IterationDescription const &)>;
struct IterationDescription {
  static IterationDescriptor descriptor;
  size_t current_iteration;
  size_t total_iterations;
  double current_elapsed_time;
};

void register_iteration_descriptor(IterationDescriptor);

}  // namespace atrip
