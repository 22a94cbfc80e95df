/* TEST INFRASTRUCTURE ONLY -- the oracle.
 *
 * A plain-C, single-threaded restatement of the (T) hot path of
 * alejandrogallo/atrip, written from the reference sources cited on each
 * function (paths relative to /root/reference).  It is the checker for the CUDA
 * path in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg;
 * nothing in the product (atrip_b200/, include/) may load it.
 *
 * Pinning: the reference's only known answer for this path (H2O,
 * integration-tests/run-h2o.sh.in:52) needs a fixture that is fetched from the
 * network and is absent here, so the restatement is pinned against THE REFERENCE
 * ITSELF: oracle/_ref/libatrip_ref.so is the reference's unmodified sources
 * compiled in this container (oracle/Makefile), tests/test_oracle.py compares
 * every function below with it on seeded inputs, and tests/golden/ holds
 * outputs of that reference build (generator: tests/golden/make_golden.py).
 *
 * All arrays are column-major (first index fastest), as CTF read_all / slice
 * produce them.
 */
#include "atrip_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- inputs -- */
/* Counter-based generator shared (by specification, not by code) with the
 * device fill kernels: value = f(seed, tensor id, column-major linear index).
 * splitmix64 finaliser; 53-bit mantissa -> u in [0,1).
 * eps_i = -2 + 1.5u, eps_a = 0.5 + 3.5u (denominators in [3,18], never zero --
 * the bench's own ranges, reference bench/main.cxx:207-222, cross zero);
 * every other tensor = scale * (u - 0.5). */
static uint64_t sm64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

double oracle_synth(uint64_t seed, int tensor_id, uint64_t idx, double scale) {
  const uint64_t key = sm64(seed ^ sm64((uint64_t)tensor_id));
  const double u = (double)(sm64(key + idx) >> 11) * (1.0 / 9007199254740992.0);
  if (tensor_id == ORACLE_EPS_I) return -2.0 + 1.5 * u;
  if (tensor_id == ORACLE_EPS_A) return 0.5 + 3.5 * u;
  return scale * (u - 0.5);
}

void oracle_fill(uint64_t seed, int tensor_id, double scale, uint64_t first,
                 uint64_t count, double *out) {
  for (uint64_t i = 0; i < count; i++)
    out[i] = oracle_synth(seed, tensor_id, first + i, scale);
}

/* ---------------------------------------------------------------- slices -- */
/* TAPHH, Unions.hpp:77-113: box {x,0,0,0}..{x+1,Nv,No,No} of Tabij[Nv,Nv,No,No]
 * -> TX[E + p Nv + q Nv No] = Tabij[x, E, p, q] */
void oracle_slice_TA(long No, long Nv, const double *Tabij, long x, double *out) {
  for (long q = 0; q < No; q++)
    for (long p = 0; p < No; p++)
      for (long E = 0; E < Nv; E++)
        out[E + p * Nv + q * Nv * No] =
            Tabij[x + E * Nv + p * Nv * Nv + q * Nv * Nv * No];
}

/* HHHA, Unions.hpp:115-152: box {0,0,0,x}..{No,No,No,x+1} of Vijka[No,No,No,Nv]
 * -> HX[p + q No + L No^2] = Vijka[p, q, L, x] */
void oracle_slice_HHHA(long No, long Nv, const double *Vijka, long x, double *out) {
  (void)Nv;
  memcpy(out, Vijka + x * No * No * No, sizeof(double) * No * No * No);
}

/* ABPH, Unions.hpp:154-197: el = x + y Nv; box {x,y,0,0}..{x+1,y+1,Nv,No} of
 * Vabci[Nv,Nv,Nv,No] -> VXY[E + r Nv] = Vabci[x, y, E, r] */
void oracle_slice_ABPH(long No, long Nv, const double *Vabci, long x, long y, double *out) {
  for (long r = 0; r < No; r++)
    for (long E = 0; E < Nv; E++)
      out[E + r * Nv] = Vabci[x + y * Nv + E * Nv * Nv + r * Nv * Nv * Nv];
}

/* ABHH / TABHH, Unions.hpp:199-278: box {x,y,0,0}..{x+1,y+1,No,No} of a
 * [Nv,Nv,No,No] tensor -> XY[p + q No] = X[x, y, p, q] */
void oracle_slice_ABHH(long No, long Nv, const double *Vabij, long x, long y, double *out) {
  for (long q = 0; q < No; q++)
    for (long p = 0; p < No; p++)
      out[p + q * No] = Vabij[x + y * Nv + p * Nv * Nv + q * Nv * Nv * No];
}

/* The 18 slices of one tuple straight from the counter-based generator, without the full
 * tensors (which are 25 GB at the bench size): same element <-> source-index maps as the
 * oracle_slice_* functions above.  out[] order = reference argument order of
 * doubles_contribution (Equations.hpp:64-96) followed by VABij, VACij, VBCij:
 * VAB VAC VBC VBA VCA VCB | HA HB HC | TA TB TC | TAB TAC TBC | VABij VACij VBCij */
void oracle_synth_tuple_slices(uint64_t seed, double scale, long No, long Nv, long a,
                               long b, long c, double **out) {
  const long abc[3] = {a, b, c};
  const long ph[6][2] = {{a, b}, {a, c}, {b, c}, {b, a}, {c, a}, {c, b}};
  const long hh[3][2] = {{a, b}, {a, c}, {b, c}};
  const uint64_t uNv = (uint64_t)Nv, uNo = (uint64_t)No;
  for (int s = 0; s < 6; s++)
    for (uint64_t r = 0; r < uNo; r++)
      for (uint64_t E = 0; E < uNv; E++)
        out[s][E + r * uNv] = oracle_synth(seed, ORACLE_VABCI,
            (uint64_t)ph[s][0] + (uint64_t)ph[s][1] * uNv + E * uNv * uNv + r * uNv * uNv * uNv, scale);
  for (int s = 0; s < 3; s++)
    oracle_fill(seed, ORACLE_VIJKA, scale, (uint64_t)abc[s] * uNo * uNo * uNo, uNo * uNo * uNo, out[6 + s]);
  for (int s = 0; s < 3; s++)
    for (uint64_t q = 0; q < uNo; q++)
      for (uint64_t p = 0; p < uNo; p++)
        for (uint64_t E = 0; E < uNv; E++)
          out[9 + s][E + p * uNv + q * uNv * uNo] = oracle_synth(seed, ORACLE_TABIJ,
              (uint64_t)abc[s] + E * uNv + p * uNv * uNv + q * uNv * uNv * uNo, scale);
  for (int s = 0; s < 3; s++)
    for (uint64_t q = 0; q < uNo; q++)
      for (uint64_t p = 0; p < uNo; p++) {
        const uint64_t idx = (uint64_t)hh[s][0] + (uint64_t)hh[s][1] * uNv + p * uNv * uNv + q * uNv * uNv * uNo;
        out[12 + s][p + q * uNo] = oracle_synth(seed, ORACLE_TABIJ, idx, scale);
        out[15 + s][p + q * uNo] = oracle_synth(seed, ORACLE_VABIJ, idx, scale);
      }
}

/* ------------------------------------------------------------- equations -- */
/* doubles_contribution, Equations.cxx:455-728.  Stated term by term as the
 * reference's own element-wise form (Equations.cxx:687-726), which the dgemm
 * path (:620-680: GEMM into _t_buffer, then reorder<perm> accumulate,
 * :26-83) evaluates to the same Tijk:
 *   holes     - sum_L  TABhh[L,j] VhhhC[i,k,L] + TABhh[i,L] VhhhC[j,k,L]
 *                    + TAChh[L,k] VhhhB[i,j,L] + TAChh[i,L] VhhhB[k,j,L]
 *                    + TBChh[L,k] VhhhA[j,i,L] + TBChh[j,L] VhhhA[k,i,L]
 *   particles + sum_E  TAphh[E,i,j] VBCph[E,k] + TAphh[E,i,k] VCBph[E,j]
 *                    + TCphh[E,k,i] VABph[E,j] + TCphh[E,k,j] VBAph[E,i]
 *                    + TBphh[E,j,i] VACph[E,k] + TBphh[E,j,k] VCAph[E,i]   */
void oracle_doubles(long No, long Nv, const double *VAB, const double *VAC,
                    const double *VBC, const double *VBA, const double *VCA,
                    const double *VCB, const double *HA, const double *HB,
                    const double *HC, const double *TA, const double *TB,
                    const double *TC, const double *TAB, const double *TAC,
                    const double *TBC, double *Tijk) {
  const long NoNo = No * No, NoNv = No * Nv;
  for (long k = 0; k < No; k++)
    for (long j = 0; j < No; j++)
      for (long i = 0; i < No; i++) {
        double t = 0.0;
        for (long L = 0; L < No; L++) {
          t -= TAB[L + j * No] * HC[i + k * No + L * NoNo];
          t -= TAB[i + L * No] * HC[j + k * No + L * NoNo];
          t -= TAC[L + k * No] * HB[i + j * No + L * NoNo];
          t -= TAC[i + L * No] * HB[k + j * No + L * NoNo];
          t -= TBC[L + k * No] * HA[j + i * No + L * NoNo];
          t -= TBC[j + L * No] * HA[k + i * No + L * NoNo];
        }
        for (long E = 0; E < Nv; E++) {
          t += TA[E + i * Nv + j * NoNv] * VBC[E + k * Nv];
          t += TA[E + i * Nv + k * NoNv] * VCB[E + j * Nv];
          t += TC[E + k * Nv + i * NoNv] * VAB[E + j * Nv];
          t += TC[E + k * Nv + j * NoNv] * VBA[E + i * Nv];
          t += TB[E + j * Nv + i * NoNv] * VAC[E + k * Nv];
          t += TB[E + j * Nv + k * NoNv] * VCA[E + i * Nv];
        }
        Tijk[i + j * No + k * NoNo] = t;
      }
}

/* singles_contribution, Equations.cxx:387-426 (Zijk must already hold Tijk:
 * the caller copies it first, Atrip.cxx:899-906) */
void oracle_singles(long No, long Nv, long a, long b, long c, const double *Tph,
                    const double *VABij, const double *VACij,
                    const double *VBCij, double *Zijk) {
  for (long k = 0; k < No; k++)
    for (long i = 0; i < No; i++)
      for (long j = 0; j < No; j++) {
        const long ijk = i + j * No + k * No * No;
        Zijk[ijk] += Tph[a + i * Nv] * VBCij[j + k * No];
        Zijk[ijk] += Tph[b + j * Nv] * VACij[i + k * No];
        Zijk[ijk] += Tph[c + k * Nv] * VABij[i + j * No];
      }
}

/* get_energy_distinct, Equations.cxx:101-180: triangle k <= j <= i with weights
 * facjk, facij; the reference's 16-blocking (:108-116) only reorders the sum and
 * is kept so that the accumulation order, hence the rounding, is the same. */
double oracle_energy_distinct(double epsabc, long No, const double *epsi,
                              const double *Tijk, const double *Zijk) {
  const long bs = 16, NN = No * No;
  double energy = 0.0;
  for (long kk = 0; kk < No; kk += bs) {
    const long kend = kk + bs < No ? kk + bs : No;
    for (long jj = kk; jj < No; jj += bs) {
      const long jend = jj + bs < No ? jj + bs : No;
      for (long ii = jj; ii < No; ii += bs) {
        const long iend = ii + bs < No ? ii + bs : No;
        for (long k = kk; k < kend; k++) {
          for (long j = jj > k ? jj : k; j < jend; j++) {
            const double facjk = j == k ? 0.5 : 1.0;
            for (long i = ii > j ? ii : j; i < iend; i++) {
              const double facij = i == j ? 0.5 : 1.0;
              const double den = epsabc - ((epsi[i] + epsi[j]) + epsi[k]);
              const double U = Zijk[i + No * j + NN * k], V = Zijk[i + No * k + NN * j],
                           W = Zijk[j + No * i + NN * k], X = Zijk[j + No * k + NN * i],
                           Y = Zijk[k + No * i + NN * j], Z = Zijk[k + No * j + NN * i];
              const double A = Tijk[i + No * j + NN * k], B = Tijk[i + No * k + NN * j],
                           C = Tijk[j + No * i + NN * k], D = Tijk[j + No * k + NN * i],
                           E = Tijk[k + No * i + NN * j], F = Tijk[k + No * j + NN * i];
              const double UXY = U + (X + Y), VWZ = V + (W + Z);
              const double ADE = A + (D + E), BCF = B + (C + F);
              const double first =
                  A * U + (B * V + (C * W + (D * X + (E * Y + F * Z))));
              const double second = (UXY - 2.0 * VWZ) * ADE;
              const double third = (VWZ - 2.0 * UXY) * BCF;
              const double value = 3.0 * first + (second + third);
              energy += ((2.0 * value) / den) * (facjk * facij);
            }
          }
        }
      }
    }
  }
  return energy;
}

/* get_energy_same, Equations.cxx:182-238: only the cyclic permutations */
double oracle_energy_same(double epsabc, long No, const double *epsi,
                          const double *Tijk, const double *Zijk) {
  const long bs = 16, NN = No * No;
  double energy = 0.0;
  for (long kk = 0; kk < No; kk += bs) {
    const long kend = kk + bs < No ? kk + bs : No;
    for (long jj = kk; jj < No; jj += bs) {
      const long jend = jj + bs < No ? jj + bs : No;
      for (long ii = jj; ii < No; ii += bs) {
        const long iend = ii + bs < No ? ii + bs : No;
        for (long k = kk; k < kend; k++) {
          for (long j = jj > k ? jj : k; j < jend; j++) {
            const double facjk = j == k ? 0.5 : 1.0;
            for (long i = ii > j ? ii : j; i < iend; i++) {
              const double facij = i == j ? 0.5 : 1.0;
              const double den = epsabc - ((epsi[i] + epsi[j]) + epsi[k]);
              const double U = Zijk[i + No * j + NN * k], V = Zijk[j + No * k + NN * i],
                           W = Zijk[k + No * i + NN * j];
              const double A = Tijk[i + No * j + NN * k], B = Tijk[j + No * k + NN * i],
                           C = Tijk[k + No * i + NN * j];
              const double ABC = A + (B + C), UVW = U + (V + W);
              const double value = 3.0 * ((A * U + B * V) + C * W) - ABC * UVW;
              energy += ((2.0 * value) / den) * (facjk * facij);
            }
          }
        }
      }
    }
  }
  return energy;
}

/* One iteration of the main loop, Atrip.cxx:855-925 (+ :928-963 for cT):
 * slices by Slice.cxx:36-50 / Atrip.cxx:861-878, 919-921; epsabc and the
 * distinct/same choice by Atrip.cxx:640-650. */
double oracle_tuple_energy(long No, long Nv, const double *epsi, const double *epsa,
                           const double *Tai, const double *Tabij,
                           const double *Vabij, const double *Vijka,
                           const double *Vabci, const double *Jijka,
                           const double *Jabci, long a, long b, long c,
                           double *Tijk_out, double *Zijk_out, double *ct) {
  const long N3 = No * No * No, NvNo = Nv * No, NN = No * No;
  double *buf = (double *)malloc(sizeof(double) *
                                 (6 * NvNo + 3 * N3 + 3 * NvNo * No + 6 * NN + 2 * N3));
  double *VAB = buf, *VAC = VAB + NvNo, *VBC = VAC + NvNo, *VBA = VBC + NvNo,
         *VCA = VBA + NvNo, *VCB = VCA + NvNo;
  double *HA = VCB + NvNo, *HB = HA + N3, *HC = HB + N3;
  double *TA = HC + N3, *TB = TA + NvNo * No, *TC = TB + NvNo * No;
  double *TAB = TC + NvNo * No, *TAC = TAB + NN, *TBC = TAC + NN;
  double *VABij = TBC + NN, *VACij = VABij + NN, *VBCij = VACij + NN;
  double *Tijk = VBCij + NN, *Zijk = Tijk + N3;

  oracle_slice_ABPH(No, Nv, Vabci, a, b, VAB);
  oracle_slice_ABPH(No, Nv, Vabci, a, c, VAC);
  oracle_slice_ABPH(No, Nv, Vabci, b, c, VBC);
  oracle_slice_ABPH(No, Nv, Vabci, b, a, VBA);
  oracle_slice_ABPH(No, Nv, Vabci, c, a, VCA);
  oracle_slice_ABPH(No, Nv, Vabci, c, b, VCB);
  oracle_slice_HHHA(No, Nv, Vijka, a, HA);
  oracle_slice_HHHA(No, Nv, Vijka, b, HB);
  oracle_slice_HHHA(No, Nv, Vijka, c, HC);
  oracle_slice_TA(No, Nv, Tabij, a, TA);
  oracle_slice_TA(No, Nv, Tabij, b, TB);
  oracle_slice_TA(No, Nv, Tabij, c, TC);
  oracle_slice_ABHH(No, Nv, Tabij, a, b, TAB);
  oracle_slice_ABHH(No, Nv, Tabij, a, c, TAC);
  oracle_slice_ABHH(No, Nv, Tabij, b, c, TBC);
  oracle_slice_ABHH(No, Nv, Vabij, a, b, VABij);
  oracle_slice_ABHH(No, Nv, Vabij, a, c, VACij);
  oracle_slice_ABHH(No, Nv, Vabij, b, c, VBCij);

  oracle_doubles(No, Nv, VAB, VAC, VBC, VBA, VCA, VCB, HA, HB, HC, TA, TB, TC,
                 TAB, TAC, TBC, Tijk);
  memcpy(Zijk, Tijk, sizeof(double) * N3);
  oracle_singles(No, Nv, a, b, c, Tai, VABij, VACij, VBCij, Zijk);

  const double epsabc = epsa[a] + epsa[b] + epsa[c];
  const int same = (a == b) != (b == c); /* Atrip.cxx:640-642 */
  double e = same ? oracle_energy_same(epsabc, No, epsi, Tijk, Zijk)
                  : oracle_energy_distinct(epsabc, No, epsi, Tijk, Zijk);
  if (Tijk_out) memcpy(Tijk_out, Tijk, sizeof(double) * N3);
  if (Zijk_out) memcpy(Zijk_out, Zijk, sizeof(double) * N3);

  if (ct) {
    *ct = e; /* without J the reference just evaluates the energy twice */
    if (Jijka && Jabci) {
      /* second doubles pass with J slices, T unchanged, Zijk kept from the V
       * pass (Atrip.cxx:928-963) */
      oracle_slice_ABPH(No, Nv, Jabci, a, b, VAB);
      oracle_slice_ABPH(No, Nv, Jabci, a, c, VAC);
      oracle_slice_ABPH(No, Nv, Jabci, b, c, VBC);
      oracle_slice_ABPH(No, Nv, Jabci, b, a, VBA);
      oracle_slice_ABPH(No, Nv, Jabci, c, a, VCA);
      oracle_slice_ABPH(No, Nv, Jabci, c, b, VCB);
      oracle_slice_HHHA(No, Nv, Jijka, a, HA);
      oracle_slice_HHHA(No, Nv, Jijka, b, HB);
      oracle_slice_HHHA(No, Nv, Jijka, c, HC);
      oracle_doubles(No, Nv, VAB, VAC, VBC, VBA, VCA, VCB, HA, HB, HC, TA, TB,
                     TC, TAB, TAC, TBC, Tijk);
      *ct = same ? oracle_energy_same(epsabc, No, epsi, Tijk, Zijk)
                 : oracle_energy_distinct(epsabc, No, epsi, Tijk, Zijk);
    }
  }
  free(buf);
  return e;
}

/* ---------------------------------------------------------------- tuples -- */
/* Tuples.cxx:91-93, 122-134 */
long oracle_n_tuples(long Nv) { return Nv * (Nv + 1) * (Nv + 2) / 6 - Nv; }

long oracle_all_tuples(long Nv, uint64_t *out, long cap) {
  long u = 0;
  for (long a = 0; a < Nv; a++)
    for (long b = a; b < Nv; b++)
      for (long c = b; c < Nv; c++) {
        if (a == b && b == c) continue;
        if (u < cap) {
          out[3 * u] = (uint64_t)a;
          out[3 * u + 1] = (uint64_t)b;
          out[3 * u + 2] = (uint64_t)c;
        }
        u++;
      }
  return u;
}

static int cmp_tuple(const void *x, const void *y) {
  const uint64_t *p = (const uint64_t *)x, *q = (const uint64_t *)y;
  for (int d = 0; d < 3; d++)
    if (p[d] != q[d]) return p[d] < q[d] ? -1 : 1;
  return 0;
}

/* sorted distinct nodes of a tuple: get_tuple_nodes + unique, Tuples.cxx:6-13,
 * 145-154 (node of an index = index % n_nodes) */
static int tuple_nodes(const uint64_t *t, long n, long nodes[3]) {
  long v[3] = {(long)(t[0] % n), (long)(t[1] % n), (long)(t[2] % n)};
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 3; j++)
      if (v[j] < v[i]) {
        long s = v[i];
        v[i] = v[j];
        v[j] = s;
      }
  int m = 0;
  for (int i = 0; i < 3; i++)
    if (m == 0 || v[i] != nodes[m - 1]) nodes[m++] = v[i];
  return m;
}

/* group_and_sort::special_distribution, Tuples.cxx:156-308.  Containers keyed
 * by the sorted node set; a 1-node container goes whole to its node, a 2-node
 * container is cut [0,half) | [half,size) between (smaller, larger) node id, a
 * 3-node container in thirds; then the "home elements fastest" swap (:267-286),
 * lexicographic sort (:290) and restoring a<=b<=c (:294). */
long oracle_group_and_sort(long n_nodes, long node_id, long Nv, uint64_t *out, long cap) {
  const long n = n_nodes, me = node_id;
  const long ntot = oracle_n_tuples(Nv);
  uint64_t *all = (uint64_t *)malloc(sizeof(uint64_t) * 3 * (ntot > 0 ? ntot : 1));
  oracle_all_tuples(Nv, all, ntot);
  const long nkeys = n * n * n;
  long *size = (long *)calloc((size_t)nkeys, sizeof(long));
  long *seen = (long *)calloc((size_t)nkeys, sizeof(long));
  /* pass 1: container sizes.  key = n0 + n1 n + n2 n^2 over the sorted nodes,
   * unused slots = the last node (so 1-d and 2-d keys stay distinct) */
  for (long t = 0; t < ntot; t++) {
    long nd[3];
    const int m = tuple_nodes(all + 3 * t, n, nd);
    const long key = nd[0] + (m > 1 ? nd[1] : nd[0]) * n + (m > 2 ? nd[2] : nd[m - 1]) * n * n;
    size[key]++;
  }
  /* pass 2: take my share of each container, in enumeration order */
  uint64_t *mine = (uint64_t *)malloc(sizeof(uint64_t) * 3 * (ntot > 0 ? ntot : 1));
  long nm = 0;
  for (long t = 0; t < ntot; t++) {
    long nd[3];
    const uint64_t *tp = all + 3 * t;
    const int m = tuple_nodes(tp, n, nd);
    const long key = nd[0] + (m > 1 ? nd[1] : nd[0]) * n + (m > 2 ? nd[2] : nd[m - 1]) * n * n;
    const long pos = seen[key]++, sz = size[key];
    int take = 0;
    if (m == 1) {
      take = nd[0] == me;
    } else if (m == 2) {
      const long half = sz / 2;
      if (me == nd[0]) take = pos < half;
      else if (me == nd[1]) take = pos >= half;
    } else {
      const long third = sz / 3;
      if (me == nd[0]) take = pos < third;
      else if (me == nd[1]) take = pos >= third && pos < 2 * third;
      else if (me == nd[2]) take = pos >= 2 * third;
    }
    if (take) {
      memcpy(mine + 3 * nm, tp, 3 * sizeof(uint64_t));
      nm++;
    }
  }
  /* home elements to the back so that non-home indices vary slowest */
  for (long t = 0; t < nm; t++) {
    uint64_t *nt = mine + 3 * t;
    const int h0 = (long)(nt[0] % n) == me, h1 = (long)(nt[1] % n) == me,
              h2 = (long)(nt[2] % n) == me;
    if (h0) {
      if (!h2) {
        uint64_t x = nt[0];
        nt[0] = nt[2];
        nt[2] = x;
      } else if (!h1) {
        uint64_t x = nt[0];
        nt[0] = nt[1];
        nt[1] = x;
      }
    } else if (h1 && !h2) {
      uint64_t x = nt[1];
      nt[1] = nt[2];
      nt[2] = x;
    }
  }
  qsort(mine, (size_t)nm, 3 * sizeof(uint64_t), cmp_tuple);
  for (long t = 0; t < nm; t++) {
    uint64_t *v = mine + 3 * t;
    for (int i = 0; i < 3; i++)
      for (int j = i + 1; j < 3; j++)
        if (v[j] < v[i]) {
          uint64_t s = v[i];
          v[i] = v[j];
          v[j] = s;
        }
  }
  for (long t = 0; t < nm && t < cap; t++) memcpy(out + 3 * t, mine + 3 * t, 3 * sizeof(uint64_t));
  free(all);
  free(size);
  free(seen);
  free(mine);
  return nm;
}

/* RankMap<F>::find, RankMap.cxx:35-85, in the one-rank-per-node layout used
 * here (ranks_per_node == 1, so node-then-local-rank round robin and
 * rank_round_robin coincide): owner = index % n_ranks, index = t0 + t1 * Nv for
 * pair slices (RankMap.cxx:43-44), t0 for single-index slices. */
long oracle_owner_single(long x, long n_ranks) { return x % n_ranks; }
long oracle_owner_pair(long x, long y, long Nv, long n_ranks) {
  return (x + y * Nv) % n_ranks;
}

/* ------------------------------------------------------------------- run -- */
/* Atrip::run, Atrip.cxx:686-1057 + 1094-1111: energy = - sum over tuples, in
 * list order in one double. */
int oracle_run(long No, long Nv, const double *epsi, const double *epsa,
               const double *Tai, const double *Tabij, const double *Vabij,
               const double *Vijka, const double *Vabci, const double *Jijka,
               const double *Jabci, const uint64_t *tuples, long n_tuples,
               double *energy, double *ct_energy) {
  uint64_t *own = NULL;
  if (!tuples) {
    n_tuples = oracle_n_tuples(Nv);
    own = (uint64_t *)malloc(sizeof(uint64_t) * 3 * (n_tuples > 0 ? n_tuples : 1));
    oracle_all_tuples(Nv, own, n_tuples);
    tuples = own;
  }
  double e = 0.0, ect = 0.0;
  for (long t = 0; t < n_tuples; t++) {
    const long a = (long)tuples[3 * t], b = (long)tuples[3 * t + 1], c = (long)tuples[3 * t + 2];
    if (a == 0 && b == 0 && c == 0) continue; /* FAKE_TUPLE, Tuples.hpp:43 */
    double ct = 0.0;
    e += oracle_tuple_energy(No, Nv, epsi, epsa, Tai, Tabij, Vabij, Vijka, Vabci,
                             Jijka, Jabci, a, b, c, NULL, NULL, &ct);
    ect += ct;
  }
  *energy = -e;
  *ct_energy = -ect;
  free(own);
  return 0;
}
