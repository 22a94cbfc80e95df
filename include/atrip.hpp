// atrip.hpp -- umbrella header of the B200 build of atrip's public API
// (same role as the reference's src/atrip.hpp:16-19).
#pragma once
#include <atrip/Atrip.hpp>
