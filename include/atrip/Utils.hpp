// Small utilities of the public API: the CTF include and the _FORMAT helper the reference's
// driver relies on (reference src/atrip/Utils.hpp:27-47, 71-77).
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include <ctf.hpp>

#include <atrip/Debug.hpp>

#ifndef _FORMAT
#  define _FORMAT(_fmt, ...)                                      \
    ([&](void) -> std::string {                                   \
      int _n = std::snprintf(nullptr, 0, _fmt, __VA_ARGS__);      \
      std::vector<char> _buf((size_t)_n + 1);                     \
      std::snprintf(_buf.data(), _buf.size(), _fmt, __VA_ARGS__); \
      return std::string(_buf.data());                            \
    })()
#endif
