// Host-side tuple lists of the engine (no CUDA here; unit-tested on CPU).
//
// Replaces TuplesDistribution::get_tuples (reference Tuples.cxx:89-141, 156-407) with one GPU per
// "node".  Unlike the reference, no rank materialises the O(Nv^3) global list: the distribution
// is computed in two streaming passes over the enumeration (container sizes, then this rank's
// share), which yields exactly the reference's lists (tests/test_tuples.py compares them with the
// reference's own special_distribution).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <vector>

namespace ab {

using Tuple = std::array<uint64_t, 3>;

inline uint64_t n_tuples_total(uint64_t Nv) { return Nv * (Nv + 1) * (Nv + 2) / 6 - Nv; }

// NAIVE: contiguous chunk of the lexicographic enumeration, padded with FAKE (Tuples.cxx:89-120)
inline std::vector<Tuple> naive_tuples(uint64_t Nv, uint64_t rank, uint64_t np) {
  const uint64_t n = n_tuples_total(Nv), per = n / np + (n % np != 0), start = per * rank, end = per * (rank + 1);
  std::vector<Tuple> out(per, Tuple{0, 0, 0});
  uint64_t g = 0, r = 0;
  for (uint64_t a = 0; a < Nv; a++)
    for (uint64_t b = a; b < Nv; b++) {
      // c runs over [b, Nv) minus the a==b==c point; skip whole rows outside [start,end) quickly
      const uint64_t row = (Nv - b) - (a == b ? 1 : 0);
      if (g + row <= start || g >= end) { g += row; continue; }
      for (uint64_t c = b; c < Nv; c++) {
        if (a == b && b == c) continue;
        if (start <= g && g < end) out[r++] = Tuple{a, b, c};
        g++;
      }
    }
  return out;
}

namespace gs {
// container key of a tuple = its sorted set of distinct home nodes (index % n), Tuples.cxx:145-154
struct Key {
  int m;
  uint32_t nd[3];
  uint64_t id(uint64_t n) const {
    return nd[0] + (m > 1 ? nd[1] : nd[0]) * n + (m > 2 ? nd[2] : nd[m - 1]) * n * n;
  }
};
inline Key key_of(uint64_t a, uint64_t b, uint64_t c, uint64_t n) {
  uint32_t v[3] = {(uint32_t)(a % n), (uint32_t)(b % n), (uint32_t)(c % n)};
  if (v[1] < v[0]) std::swap(v[0], v[1]);
  if (v[2] < v[1]) std::swap(v[1], v[2]);
  if (v[1] < v[0]) std::swap(v[0], v[1]);
  Key k{0, {0, 0, 0}};
  for (int i = 0; i < 3; i++)
    if (k.m == 0 || v[i] != k.nd[k.m - 1]) k.nd[k.m++] = v[i];
  return k;
}
// does node `me` take position pos of a container of size sz?  (Tuples.cxx:188-258: whole /
// halves / thirds, the last owner takes the remainder)
inline bool takes(const Key &k, uint64_t me, uint64_t pos, uint64_t sz) {
  if (k.m == 1) return k.nd[0] == me;
  if (k.m == 2) {
    const uint64_t half = sz / 2;
    if (me == k.nd[0]) return pos < half;
    if (me == k.nd[1]) return pos >= half;
    return false;
  }
  const uint64_t third = sz / 3;
  if (me == k.nd[0]) return pos < third;
  if (me == k.nd[1]) return pos >= third && pos < 2 * third;
  if (me == k.nd[2]) return pos >= 2 * third;
  return false;
}
}  // namespace gs

// GROUP_AND_SORT for node `me` of `n` (Tuples.cxx:156-308), padded with FAKE to the longest
// node's list (Tuples.cxx:346-377).  If counts != nullptr it receives every node's real count.
inline std::vector<Tuple> group_and_sort_tuples(uint64_t Nv, uint64_t me, uint64_t n, bool pad = true,
                                                std::vector<uint64_t> *counts = nullptr) {
  const uint64_t nkeys = n * n * n;
  std::vector<uint64_t> size(nkeys, 0), seen(nkeys, 0);
  for (uint64_t a = 0; a < Nv; a++)
    for (uint64_t b = a; b < Nv; b++)
      for (uint64_t c = b; c < Nv; c++) {
        if (a == b && b == c) continue;
        size[gs::key_of(a, b, c, n).id(n)]++;
      }
  // every node's count follows from the container sizes alone
  std::vector<uint64_t> cnt(n, 0);
  for (uint64_t x = 0; x < n; x++) {
    cnt[x] += size[x + x * n + x * n * n];  // one home node: the whole container
    for (uint64_t y = x + 1; y < n; y++) {
      const uint64_t s2 = size[x + y * n + y * n * n];  // two home nodes: halves
      cnt[x] += s2 / 2;
      cnt[y] += s2 - s2 / 2;
      for (uint64_t z = y + 1; z < n; z++) {
        const uint64_t s3 = size[x + y * n + z * n * n];  // three: thirds
        cnt[x] += s3 / 3;
        cnt[y] += s3 / 3;
        cnt[z] += s3 - 2 * (s3 / 3);
      }
    }
  }
  if (counts) *counts = cnt;
  std::vector<Tuple> mine;
  mine.reserve(cnt[me]);
  for (uint64_t a = 0; a < Nv; a++)
    for (uint64_t b = a; b < Nv; b++)
      for (uint64_t c = b; c < Nv; c++) {
        if (a == b && b == c) continue;
        const gs::Key k = gs::key_of(a, b, c, n);
        const uint64_t id = k.id(n), pos = seen[id]++;
        if (gs::takes(k, me, pos, size[id])) mine.push_back(Tuple{a, b, c});
      }
  // home elements to the back so the non-home indices vary slowest after sorting (:267-286)
  for (auto &t : mine) {
    const bool h0 = t[0] % n == me, h1 = t[1] % n == me, h2 = t[2] % n == me;
    if (h0) {
      if (!h2) std::swap(t[0], t[2]);
      else if (!h1) std::swap(t[0], t[1]);
    } else if (h1 && !h2) std::swap(t[1], t[2]);
  }
  std::sort(mine.begin(), mine.end());
  for (auto &t : mine) std::sort(t.begin(), t.end());
  if (pad) {
    const uint64_t mx = *std::max_element(cnt.begin(), cnt.end());
    mine.resize(mx, Tuple{0, 0, 0});
  }
  return mine;
}

}  // namespace ab
