// Host-side tuple lists of the engine (no CUDA here; unit-tested on CPU).
//
// Replaces TuplesDistribution::get_tuples (reference Tuples.cxx:89-141, 156-407) with one GPU per
// "node".  Unlike the reference, no rank materialises (or even walks) the O(Nv^3) global list: the
// container sizes are closed-form per residue class and a rank only visits the tuples that can be
// its own (see group_and_sort_tuples), which yields exactly the reference's lists (tests/test_host.py
// compares them with the reference's own special_distribution).
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

namespace gs {
// pass 1 of group-and-sort: container sizes (per sorted node set) and every node's tuple count
struct Census {
  std::vector<Key> keys;        // key of every residue triple (ra, rb, rc)
  std::vector<uint32_t> kid;    // its container id
  std::vector<uint64_t> size;   // tuples per container
  std::vector<uint64_t> cnt;    // tuples per node
};
inline Census census(uint64_t Nv, uint64_t n) {
  Census C;
  const uint64_t nkeys = n * n * n;
  // key of every residue triple (ra, rb, rc)
  std::vector<Key> &keys = C.keys;
  std::vector<uint32_t> &kid = C.kid;
  keys.resize(nkeys);
  kid.resize(nkeys);
  for (uint64_t ra = 0; ra < n; ra++)
    for (uint64_t rb = 0; rb < n; rb++)
      for (uint64_t rc = 0; rc < n; rc++) {
        const Key k = key_of(ra, rb, rc, n);
        keys[(ra * n + rb) * n + rc] = k;
        kid[(ra * n + rb) * n + rc] = (uint32_t)k.id(n);
      }
  // number of c in [lo, Nv) with c % n == r
  auto count_c = [&](uint64_t lo, uint64_t r) -> uint64_t {
    const uint64_t first = lo + (r + n - lo % n) % n;
    return first < Nv ? (Nv - 1 - first) / n + 1 : 0;
  };
  std::vector<uint64_t> &size = C.size;
  size.assign(nkeys, 0);
  for (uint64_t a = 0; a < Nv; a++)
    for (uint64_t b = a; b < Nv; b++) {
      const uint64_t lo = b + (a == b ? 1 : 0);  // a == b == c is not a tuple
      const uint64_t base = ((a % n) * n + b % n) * n;
      for (uint64_t r = 0; r < n; r++) size[kid[base + r]] += count_c(lo, r);
    }
  // every node's count follows from the container sizes alone
  std::vector<uint64_t> &cnt = C.cnt;
  cnt.assign(n, 0);
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
  return C;
}
}  // namespace gs

// GROUP_AND_SORT for node `me` of `n` (Tuples.cxx:156-308), padded with FAKE to the longest
// node's list (Tuples.cxx:346-377).  If counts != nullptr it receives every node's real count.
//
// The reference walks the whole O(Nv^3) enumeration on every node leader.  Here
//   * the container sizes come from residue-class counts (O(Nv^2 n) instead of O(Nv^3)): for a
//     pair (a,b) the c's of each residue class form an arithmetic progression whose length is
//     closed-form;
//   * the second pass visits only the tuples that have an index congruent to `me` -- a node only
//     ever takes from containers whose node set contains it, and every tuple of such a container
//     has such an index -- in the reference's lexicographic order, so the positions inside the
//     containers (which decide the halves / thirds) are the reference's: (a,b) pairs with a home
//     index walk all c, the others only c = me (mod n);
//   * the sort runs on packed 64-bit keys (21 bits per index).
// The lists are the reference's, tuple for tuple (tests/test_host.py compares hashes with the
// reference's own special_distribution and with the oracle).
inline std::vector<Tuple> group_and_sort_tuples(uint64_t Nv, uint64_t me, uint64_t n, bool pad = true,
                                                std::vector<uint64_t> *counts = nullptr) {
  const gs::Census C = gs::census(Nv, n);
  const std::vector<gs::Key> &keys = C.keys;
  const std::vector<uint32_t> &kid = C.kid;
  const std::vector<uint64_t> &size = C.size, &cnt = C.cnt;
  std::vector<uint64_t> seen(n * n * n, 0);
  if (counts) *counts = cnt;
  // my share, with the home elements moved to the back so that the non-home indices vary slowest
  // after sorting (:267-286); packed as t0 << 42 | t1 << 21 | t2 for the sort
  std::vector<uint64_t> packed;
  packed.reserve(cnt[me]);
  auto visit = [&](uint64_t a, uint64_t b, uint64_t c, uint64_t base) {
    const uint64_t rc = c % n, slot = base + rc, id = kid[slot], pos = seen[id]++;
    if (!gs::takes(keys[slot], me, pos, size[id])) return;
    uint64_t t[3] = {a, b, c};
    const bool h0 = a % n == me, h1 = b % n == me, h2 = rc == me;
    if (h0) {
      if (!h2) std::swap(t[0], t[2]);
      else if (!h1) std::swap(t[0], t[1]);
    } else if (h1 && !h2) std::swap(t[1], t[2]);
    packed.push_back(t[0] << 42 | t[1] << 21 | t[2]);
  };
  for (uint64_t a = 0; a < Nv; a++)
    for (uint64_t b = a; b < Nv; b++) {
      const uint64_t lo = b + (a == b ? 1 : 0), base = ((a % n) * n + b % n) * n;
      if (a % n == me || b % n == me) {
        for (uint64_t c = lo; c < Nv; c++) visit(a, b, c, base);
      } else {
        for (uint64_t c = lo + (me + n - lo % n) % n; c < Nv; c += n) visit(a, b, c, base);
      }
    }
  std::sort(packed.begin(), packed.end());
  std::vector<Tuple> mine(packed.size());
  for (size_t i = 0; i < packed.size(); i++) {
    Tuple t{packed[i] >> 42, (packed[i] >> 21) & 0x1fffff, packed[i] & 0x1fffff};
    std::sort(t.begin(), t.end());
    mine[i] = t;
  }
  if (pad) {
    const uint64_t mx = *std::max_element(cnt.begin(), cnt.end());
    mine.resize(mx, Tuple{0, 0, 0});
  }
  return mine;
}

}  // namespace ab
