// Host-side slice ownership and per-batch fetch schedule of the engine (no CUDA here; unit-tested
// on CPU, single process and 2-process gloo).
//
// Replaces, for one GPU per "node":
//   RankMap<F>::find (reference RankMap.cxx:35-85)        -> ShardMap::owner*/slot*
//   Slice::subtuple_by_slice (Slice.cxx:36-50)            -> the 12 needs of a tuple in plan_batch
//   SliceUnion::build_local_database (SliceUnion.cxx:36-171: SelfSufficient / Recycled / Fetch)
//                                                         -> plan_batch: local slot, deduplicated
//                                                            cache slot, or a fetch range
//   the per-tuple MPI_Allgather of the slice database (Atrip.cxx:443-453) + send/receive
//   (SliceUnion.cxx:365-505)                              -> one request list per peer per BATCH of
//                                                            tuples, contiguous slot ranges merged
//   clear_unused_slices_for_next_tuple (SliceUnion.cxx:173-290) -> SliceCache: slots re-assigned in
//                                                            ring order once two batches old
//
// Ownership (DESIGN.md "Multi-GPU"):
//   A  slices (TAPHH+HHHA of x)                      owner x % n               (RankMap.cxx:43-82)
//   B  slices (ABPH+TABHH of the ordered pair y,z)   owner y % n  -- the pair lives with its FIRST
//      index; equal to the reference's (y + z Nv) % n whenever Nv % n == 0
//   V  slices (ABHH of y <= z)                       stored by owner(y) AND owner(z) (they are
//      small), requests go to owner(y)
// Slots at the owner are closed-form, so a requester names remote slices by the owner's slot and
// B slices of a fixed y whose z share a residue class are consecutive: the B needs of a run of
// tuples (p0, p1, z = home, home + n, ...) merge into one range per (p0, .) / (p1, .) row.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <utility>
#include <vector>

#include "tuples.hpp"

namespace ab {

enum SliceKind : int { KA = 0, KB = 1, KV = 2 };

// One tuple of a device batch: the tuple and the store slots of its 12 slices.  A slot below the
// owned count of its store addresses the owned part, otherwise (slot - owned) addresses the cache.
struct TupleRec {
  int a, b, c, fake;
  int ax[3];   // AX slots of a, b, c
  int by[6];   // BY slots of (b,c) (a,c) (c,b)' (a,b) (c,a)' (b,a)'   (' = transposed hole part)
  int vij[3];  // VIJ slots of (b,c) (a,c) (a,b)
};
static_assert(sizeof(TupleRec) == 64, "TupleRec is 16 ints");

struct ShardMap {
  int64_t Nv = 1;
  int n = 1, me = 0;
  std::vector<int64_t> cls_off;  // position of residue class q among z = 0..Nv-1, q = 0..n

  ShardMap() : cls_off{0, 1} {}
  ShardMap(int64_t Nv_, int n_, int me_) : Nv(Nv_), n(n_), me(me_), cls_off((size_t)n_ + 1, 0) {
    for (int q = 0; q < n; q++) cls_off[q + 1] = cls_off[q] + cnt(q);
  }
  // number of indices congruent to r (mod n) below Nv
  int64_t cnt(int r) const { return (Nv - r + n - 1) / n; }

  int ownerA(int64_t x) const { return (int)(x % n); }
  int64_t slotA(int64_t x) const { return x / n; }
  int64_t ownedA(int r) const { return cnt(r); }

  // B id: y + z Nv for the ordered pair (y,z); Nv^2 + y for the transposed diagonal (y,y)'
  int64_t idB(int64_t y, int64_t z, bool transposed) const { return (y == z && transposed) ? Nv * Nv + y : y + z * Nv; }
  int ownerB(int64_t id) const { return (int)((id < Nv * Nv ? id % Nv : id - Nv * Nv) % n); }
  int64_t slotB(int64_t id) const {
    if (id >= Nv * Nv) return ((id - Nv * Nv) / n) * (Nv + 1) + Nv;
    const int64_t y = id % Nv, z = id / Nv;
    return (y / n) * (Nv + 1) + cls_off[z % n] + z / n;
  }
  int64_t ownedB(int r) const { return cnt(r) * (Nv + 1); }

  // V id: y + z Nv, y <= z.  Part 1 at owner(y): rows y' = y / n hold z = y..Nv-1.  Part 2 at
  // owner(z) (when owner(y) differs): rows z' = z / n hold the y <= z not owned by that rank.
  int ownerV(int64_t id) const { return (int)((id % Nv) % n); }
  int64_t off1(int r, int64_t t) const { return t * Nv - (int64_t)n * t * (t - 1) / 2 - (int64_t)r * t; }
  int64_t slotV1(int64_t y, int64_t z) const { return off1((int)(y % n), y / n) + (z - y); }
  int64_t nV1(int r) const { return off1(r, cnt(r)); }
  int64_t off2(int r, int64_t t) const { return (int64_t)(n - 1) * t * (t - 1) / 2 + (int64_t)r * t; }
  int64_t ownedV(int r) const { return nV1(r) + off2(r, cnt(r)); }
  // slot of V(y,z) in THIS rank's store, or -1
  int64_t localV(int64_t y, int64_t z) const {
    if (y % n == me) return slotV1(y, z);
    if (z % n == me) {
      const int64_t before = y > me ? (y - 1 - me) / n + 1 : 0;  // y'' < y owned by me
      return nV1(me) + off2(me, z / n) + (y - before);
    }
    return -1;
  }
  int64_t owned(int kind, int r) const { return kind == KA ? ownedA(r) : (kind == KB ? ownedB(r) : ownedV(r)); }
};

// contiguous slots [src_slot, src_slot + count) of `kind` at the owner -> cache slots
// [dst_slot, dst_slot + count) of the requester's fetch cache of that kind
struct FetchRange {
  int kind;
  int64_t src_slot;
  int64_t count;
  int64_t dst_slot;
};

struct BatchPlan {
  std::vector<TupleRec> recs;
  std::vector<std::vector<FetchRange>> fetch;  // [peer]
  int64_t used[3] = {0, 0, 0};                 // slices fetched for this batch per kind (cache misses)
  int64_t hits[3] = {0, 0, 0};                 // remote slices found in the cache (fetched for an earlier batch)
  bool overflow = false;                       // the cache had no free slot left: capacity too small
};

inline bool is_fake(const Tuple &t) { return t[0] == 0 && t[1] == 0 && t[2] == 0; }

// Fetch cache of one rank: cap[kind] slots per store, shared by all batches of all runs.  Plays the
// role of the reference's slice buffers with their Recycled / exact-match reuse and of
// clear_unused_slices_for_next_tuple (SliceUnion.cxx:66-137, 173-290), at batch granularity:
//   * a remote slice stays addressable (key -> slot) until its slot is handed to another slice, so
//     a batch re-uses what earlier batches fetched (the two slowly varying indices of a group-and-
//     sort run keep their A slices for the whole run);
//   * a slot may be re-assigned while planning batch `serial` only if the last batch that addressed
//     it has serial <= serial - 2: batch serial - 1 may still be computing when the copies of batch
//     `serial` are in flight (the engine orders the copies behind the reduction of serial - 2);
//   * free slots are taken in ring order, so the misses of a batch, sorted by owner slot, mostly
//     land in consecutive slots and merge into few copies.
struct SliceCache {
  static constexpr uint64_t EMPTY = ~0ull;
  int64_t cap[3] = {0, 0, 0};
  std::vector<uint64_t> key[3];   // slot -> slice held
  std::vector<int64_t> last[3];   // slot -> serial of the last batch that addressed it
  std::vector<std::pair<uint64_t, int32_t>> table[3];  // open-addressing hash: key -> slot
  int64_t hand[3] = {0, 0, 0};
  size_t touched[3] = {0, 0, 0};  // table entries that are not "never used" any more (live + tombstones)

  void reset(const int64_t cap_[3]) {
    for (int k = 0; k < 3; k++) {
      cap[k] = cap_[k];
      key[k].assign((size_t)cap[k], EMPTY);
      last[k].assign((size_t)cap[k], INT64_MIN / 2);
      size_t n = 16;
      while (n < 4 * (size_t)cap[k]) n <<= 1;
      table[k].assign(n, {EMPTY, -1});
      hand[k] = 0;
      touched[k] = 0;
    }
  }
  // the stores behind the cached copies changed: forget every slice, keep the capacity
  void invalidate() { reset(cap); }

  static size_t hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 29;
    return (size_t)k;
  }
  int64_t find(int kind, uint64_t k) const {
    const auto &t = table[kind];
    for (size_t i = hash(k) & (t.size() - 1);; i = (i + 1) & (t.size() - 1)) {
      if (t[i].first == k) return t[i].second;
      if (t[i].first == EMPTY && t[i].second == -1) return -1;  // never used: end of the probe chain
    }
  }
  void insert(int kind, uint64_t k, int32_t slot) {
    auto &t = table[kind];
    for (size_t i = hash(k) & (t.size() - 1);; i = (i + 1) & (t.size() - 1))
      if (t[i].first == EMPTY) {  // never used or tombstone
        if (t[i].second == -1) touched[kind]++;
        t[i] = {k, slot};
        return;
      }
  }
  void erase(int kind, uint64_t k) {
    auto &t = table[kind];
    for (size_t i = hash(k) & (t.size() - 1);; i = (i + 1) & (t.size() - 1)) {
      if (t[i].first == k) {
        t[i] = {EMPTY, -2};  // tombstone: keeps the probe chain intact
        return;
      }
      if (t[i].first == EMPTY && t[i].second == -1) return;
    }
  }
  // a free slot for a new slice of batch `serial` (-1: none; the cache is too small)
  int64_t take(int kind, uint64_t k, int64_t serial) {
    const int64_t n = cap[kind];
    for (int64_t step = 0; step < n; step++) {
      const int64_t s = hand[kind];
      hand[kind] = s + 1 == n ? 0 : s + 1;
      if (last[kind][(size_t)s] <= serial - 2) {
        if (key[kind][(size_t)s] != EMPTY) erase(kind, key[kind][(size_t)s]);
        key[kind][(size_t)s] = k;
        last[kind][(size_t)s] = serial;
        insert(kind, k, (int32_t)s);
        if (2 * touched[kind] > table[kind].size()) compact(kind);
        return s;
      }
    }
    return -1;
  }
  // tombstones accumulate (every re-assigned slot leaves one): rebuild the table before the probe
  // chains run out of never-used entries
  void compact(int kind) {
    auto &t = table[kind];
    for (auto &e : t) e = {EMPTY, -1};
    touched[kind] = 0;
    for (int64_t s = 0; s < cap[kind]; s++)
      if (key[kind][(size_t)s] != EMPTY) insert(kind, key[kind][(size_t)s], (int32_t)s);
  }
};

// Slots for the tuples t[0..n) on rank m.me, batch number `serial`.  owned[kind] = owned slot count
// of the store (cache slot s is addressed as owned + s).  Remote slices already in the cache --
// fetched for an earlier tuple of this batch or for an earlier batch -- are reused (the reference's
// "Recycled"/exact-match cases, SliceUnion.cxx:66-137); the rest become fetch ranges.
inline void plan_batch(const ShardMap &m, const Tuple *t, size_t n, const int64_t owned[3], SliceCache &cache,
                       int64_t serial, BatchPlan &out) {
  struct Need {
    uint64_t key;
    uint32_t rec, field;
  };
  static thread_local std::vector<Need> needs;
  static thread_local std::vector<std::pair<size_t, size_t>> misses;  // [first, last) runs of `needs` with one key
  needs.clear();
  misses.clear();
  out.recs.resize(n);
  out.fetch.assign((size_t)m.n, {});
  for (int k = 0; k < 3; k++) out.used[k] = out.hits[k] = 0;
  out.overflow = false;
  const int64_t Nv = m.Nv;
  auto remote = [&](int peer, int kind, int64_t slot, size_t rec, int field) {
    needs.push_back(Need{((uint64_t)peer << 48) | ((uint64_t)kind << 44) | (uint64_t)slot, (uint32_t)rec, (uint32_t)field});
  };
  for (size_t i = 0; i < n; i++) {
    TupleRec &r = out.recs[i];
    const int64_t abc[3] = {(int64_t)t[i][0], (int64_t)t[i][1], (int64_t)t[i][2]};
    r.a = (int)abc[0];
    r.b = (int)abc[1];
    r.c = (int)abc[2];
    r.fake = is_fake(t[i]);
    int *f = &r.ax[0];
    for (int k = 0; k < 12; k++) f[k] = 0;
    if (r.fake) continue;
    for (int k = 0; k < 3; k++) {
      const int o = m.ownerA(abc[k]);
      if (o == m.me) r.ax[k] = (int)m.slotA(abc[k]);
      else remote(o, KA, m.slotA(abc[k]), i, 4 + k);
    }
    // (y, z, transposed) per class/piece, contraction.cuh
    const int64_t yz[6][3] = {{abc[1], abc[2], 0}, {abc[0], abc[2], 0}, {abc[2], abc[1], 1},
                              {abc[0], abc[1], 0}, {abc[2], abc[0], 1}, {abc[1], abc[0], 1}};
    for (int k = 0; k < 6; k++) {
      const int64_t id = m.idB(yz[k][0], yz[k][1], yz[k][2] != 0);
      const int o = m.ownerB(id);
      if (o == m.me) r.by[k] = (int)m.slotB(id);
      else remote(o, KB, m.slotB(id), i, 7 + k);
    }
    const int64_t vp[3][2] = {{abc[1], abc[2]}, {abc[0], abc[2]}, {abc[0], abc[1]}};
    for (int k = 0; k < 3; k++) {
      const int64_t s = m.localV(vp[k][0], vp[k][1]);
      if (s >= 0) r.vij[k] = (int)s;
      else remote(m.ownerV(vp[k][0] + vp[k][1] * Nv), KV, m.slotV1(vp[k][0], vp[k][1]), i, 13 + k);
    }
  }
  if (needs.empty()) return;
  std::sort(needs.begin(), needs.end(), [](const Need &x, const Need &y) { return x.key < y.key; });
  auto kind_of = [](uint64_t key) { return (int)((key >> 44) & 15); };
  auto assign = [&](size_t lo, size_t hi, int kind, int64_t slot) {
    for (size_t q = lo; q < hi; q++) reinterpret_cast<int *>(&out.recs[needs[q].rec])[needs[q].field] = (int)(owned[kind] + slot);
  };
  // pass 1: hits keep their slot (and are marked in use before any slot is re-assigned)
  for (size_t lo = 0; lo < needs.size();) {
    size_t hi = lo + 1;
    while (hi < needs.size() && needs[hi].key == needs[lo].key) hi++;
    const int kind = kind_of(needs[lo].key);
    const int64_t s = cache.find(kind, needs[lo].key);
    if (s >= 0) {
      cache.last[kind][(size_t)s] = serial;
      out.hits[kind]++;
      assign(lo, hi, kind, s);
    } else {
      misses.push_back({lo, hi});
    }
    lo = hi;
  }
  // pass 2: misses take free slots in ring order, in owner-slot order, and merge into ranges
  for (const auto &mr : misses) {
    const uint64_t key = needs[mr.first].key;
    const int peer = (int)(key >> 48), kind = kind_of(key);
    const int64_t src = (int64_t)(key & ((1ull << 44) - 1));
    const int64_t s = cache.take(kind, key, serial);
    if (s < 0) {
      out.overflow = true;
      return;
    }
    out.used[kind]++;
    auto &fr = out.fetch[(size_t)peer];
    if (!fr.empty() && fr.back().kind == kind && fr.back().src_slot + fr.back().count == src &&
        fr.back().dst_slot + fr.back().count == s)
      fr.back().count++;
    else
      fr.push_back(FetchRange{kind, src, 1, s});
    assign(mr.first, mr.second, kind, s);
  }
}

// Cache slots needed per kind so that ANY window of `batch` consecutive tuples of the list fits
// one cache region (sliding-window count of distinct remote slices).
inline void cache_need(const ShardMap &m, const Tuple *t, size_t n, size_t batch, int64_t cap[3]) {
  cap[0] = cap[1] = cap[2] = 0;
  if (m.n == 1 || n == 0) return;
  const int64_t Nv = m.Nv;
  std::vector<int64_t> last[3];
  last[KA].assign((size_t)Nv, -1);
  last[KB].assign((size_t)(Nv * Nv + Nv), -1);
  last[KV].assign((size_t)(Nv * Nv), -1);
  int64_t cur[3] = {0, 0, 0};
  // ids of the remote needs of tuple i (global ids, not slots)
  auto ids_of = [&](const Tuple &tp, int64_t ids[12], int kinds[12]) {
    int k = 0;
    if (is_fake(tp)) return 0;
    const int64_t a = (int64_t)tp[0], b = (int64_t)tp[1], c = (int64_t)tp[2];
    const int64_t x[3] = {a, b, c};
    for (int j = 0; j < 3; j++)
      if (m.ownerA(x[j]) != m.me) { ids[k] = x[j]; kinds[k++] = KA; }
    const int64_t yz[6][3] = {{b, c, 0}, {a, c, 0}, {c, b, 1}, {a, b, 0}, {c, a, 1}, {b, a, 1}};
    for (int j = 0; j < 6; j++) {
      const int64_t id = m.idB(yz[j][0], yz[j][1], yz[j][2] != 0);
      if (m.ownerB(id) != m.me) { ids[k] = id; kinds[k++] = KB; }
    }
    const int64_t vp[3][2] = {{b, c}, {a, c}, {a, b}};
    for (int j = 0; j < 3; j++)
      if (m.localV(vp[j][0], vp[j][1]) < 0) { ids[k] = vp[j][0] + vp[j][1] * Nv; kinds[k++] = KV; }
    return k;
  };
  int64_t ids[12];
  int kinds[12];
  for (size_t i = 0; i < n; i++) {
    if (i >= batch) {  // tuple i - batch leaves the window
      const int64_t gone = (int64_t)(i - batch);
      const int k = ids_of(t[gone], ids, kinds);
      for (int j = 0; j < k; j++)
        if (last[kinds[j]][(size_t)ids[j]] == gone) {
          last[kinds[j]][(size_t)ids[j]] = -1;
          cur[kinds[j]]--;
        }
    }
    const int k = ids_of(t[i], ids, kinds);
    for (int j = 0; j < k; j++) {
      int64_t &l = last[kinds[j]][(size_t)ids[j]];
      if (l < 0) cur[kinds[j]]++;
      l = (int64_t)i;
    }
    for (int q = 0; q < 3; q++) cap[q] = std::max(cap[q], cur[q]);
  }
}

// wire format of a request list: [nranges, (kind, src_slot, count) x nranges] as int32
inline size_t request_capacity_ints(size_t batch) { return 1 + 3 * 12 * batch; }
inline void encode_requests(const std::vector<FetchRange> &fr, int32_t *buf) {
  buf[0] = (int32_t)fr.size();
  for (size_t i = 0; i < fr.size(); i++) {
    buf[1 + 3 * i] = fr[i].kind;
    buf[2 + 3 * i] = (int32_t)fr[i].src_slot;
    buf[3 + 3 * i] = (int32_t)fr[i].count;
  }
}

}  // namespace ab
