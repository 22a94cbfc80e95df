/* TEST INFRASTRUCTURE ONLY -- the oracle, std::complex<double> instantiation.
 *
 * Plain-C99 (double _Complex) restatement of what the reference computes when
 * it is instantiated with F = Complex (Atrip.cxx:1136, Equations.cxx:730-795):
 * the same hot path as atrip_oracle.c with the three conjugations the complex
 * field adds,
 *   - the hole integrals are conjugated before the hole GEMMs
 *     (MAYBE_CONJ(_vhhh, VhhhX), Equations.cxx:623-648, Operations.hpp:28-53);
 *     the particle GEMMs use "T", not "C" (Equations.cxx:547-560): no conjugate;
 *   - Tijk is conjugated inside the energy sums (Equations.cxx:135-146,
 *     207-212) and only the real part of the sum is kept (:176-178, :234-236);
 *   - epsabc is the real part of eps_a + eps_b + eps_c (Atrip.cxx:643-646),
 *     eps_i stays complex.
 * Pinned against the reference itself (oracle/_ref, ref_*_z in ref_driver.cxx)
 * by tests/test_oracle_complex.py and the "complex_*" entries of
 * tests/golden/reference_vectors.json.
 *
 * Complex arrays are interleaved (re, im) pairs = the memory layout of
 * std::complex<double>, column-major.  Synthetic complex inputs: element e of
 * tensor t is (synth(t, 2e), synth(t, 2e+1)); eps_i / eps_a are
 * (synth(t, e), 0), i.e. the real case's values with a zero imaginary part.
 */
#include <complex.h>
#include <stdlib.h>
#include <string.h>

#include "atrip_oracle.h"

typedef double _Complex zd;

void oracle_fill_z(uint64_t seed, int tensor_id, double scale, uint64_t first,
                   uint64_t count, double *out) {
  for (uint64_t i = 0; i < count; i++) {
    const uint64_t e = first + i;
    if (tensor_id == ORACLE_EPS_I || tensor_id == ORACLE_EPS_A) {
      out[2 * i] = oracle_synth(seed, tensor_id, e, scale);
      out[2 * i + 1] = 0.0;
    } else {
      out[2 * i] = oracle_synth(seed, tensor_id, 2 * e, scale);
      out[2 * i + 1] = oracle_synth(seed, tensor_id, 2 * e + 1, scale);
    }
  }
}

/* slices: same index maps as the real case (Unions.hpp:77-278) */
static void slice_TA(long No, long Nv, const zd *Tabij, long x, zd *out) {
  for (long q = 0; q < No; q++)
    for (long p = 0; p < No; p++)
      for (long E = 0; E < Nv; E++)
        out[E + p * Nv + q * Nv * No] = Tabij[x + E * Nv + p * Nv * Nv + q * Nv * Nv * No];
}
static void slice_HHHA(long No, const zd *Vijka, long x, zd *out) {
  memcpy(out, Vijka + x * No * No * No, sizeof(zd) * No * No * No);
}
static void slice_ABPH(long No, long Nv, const zd *Vabci, long x, long y, zd *out) {
  for (long r = 0; r < No; r++)
    for (long E = 0; E < Nv; E++)
      out[E + r * Nv] = Vabci[x + y * Nv + E * Nv * Nv + r * Nv * Nv * Nv];
}
static void slice_ABHH(long No, long Nv, const zd *Vabij, long x, long y, zd *out) {
  for (long q = 0; q < No; q++)
    for (long p = 0; p < No; p++)
      out[p + q * No] = Vabij[x + y * Nv + p * Nv * Nv + q * Nv * Nv * No];
}

/* doubles_contribution<Complex>: the twelve terms of the production (dgemm)
 * path, Equations.cxx:620-680, written element-wise: the hole integrals are
 * conjugated (MAYBE_CONJ(_vhhh, VhhhX), :623-648), nothing else is.  NOTE: the
 * reference's BLAS-free loop build (:685-727) does NOT conjugate them (it
 * carries a "TODO: conjugate T for complex", :695) and therefore disagrees with
 * its own dgemm build for complex input; the dgemm build (what configure
 * selects, configure.ac ATRIP_USE_DGEMM) is the one followed and pinned. */
void oracle_doubles_z(long No, long Nv, const double *VAB_, const double *VAC_,
                      const double *VBC_, const double *VBA_, const double *VCA_,
                      const double *VCB_, const double *HA_, const double *HB_,
                      const double *HC_, const double *TA_, const double *TB_,
                      const double *TC_, const double *TAB_, const double *TAC_,
                      const double *TBC_, double *Tijk_) {
  const zd *VAB = (const zd *)VAB_, *VAC = (const zd *)VAC_, *VBC = (const zd *)VBC_,
           *VBA = (const zd *)VBA_, *VCA = (const zd *)VCA_, *VCB = (const zd *)VCB_,
           *HA = (const zd *)HA_, *HB = (const zd *)HB_, *HC = (const zd *)HC_,
           *TA = (const zd *)TA_, *TB = (const zd *)TB_, *TC = (const zd *)TC_,
           *TAB = (const zd *)TAB_, *TAC = (const zd *)TAC_, *TBC = (const zd *)TBC_;
  zd *Tijk = (zd *)Tijk_;
  const long NoNo = No * No, NoNv = No * Nv;
  for (long k = 0; k < No; k++)
    for (long j = 0; j < No; j++)
      for (long i = 0; i < No; i++) {
        zd t = 0.0;
        for (long L = 0; L < No; L++) {
          t -= TAB[L + j * No] * conj(HC[i + k * No + L * NoNo]);
          t -= TAB[i + L * No] * conj(HC[j + k * No + L * NoNo]);
          t -= TAC[L + k * No] * conj(HB[i + j * No + L * NoNo]);
          t -= TAC[i + L * No] * conj(HB[k + j * No + L * NoNo]);
          t -= TBC[L + k * No] * conj(HA[j + i * No + L * NoNo]);
          t -= TBC[j + L * No] * conj(HA[k + i * No + L * NoNo]);
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

/* singles_contribution<Complex>, Equations.cxx:387-426: plain products */
void oracle_singles_z(long No, long Nv, long a, long b, long c, const double *Tph_,
                      const double *VABij_, const double *VACij_,
                      const double *VBCij_, double *Zijk_) {
  const zd *Tph = (const zd *)Tph_, *VABij = (const zd *)VABij_,
           *VACij = (const zd *)VACij_, *VBCij = (const zd *)VBCij_;
  zd *Zijk = (zd *)Zijk_;
  for (long k = 0; k < No; k++)
    for (long i = 0; i < No; i++)
      for (long j = 0; j < No; j++) {
        const long ijk = i + j * No + k * No * No;
        Zijk[ijk] += Tph[a + i * Nv] * VBCij[j + k * No];
        Zijk[ijk] += Tph[b + j * Nv] * VACij[i + k * No];
        Zijk[ijk] += Tph[c + k * Nv] * VABij[i + j * No];
      }
}

/* get_energy_distinct<Complex>, Equations.cxx:101-180 */
double oracle_energy_distinct_z(double epsabc, long No, const double *epsi_,
                                const double *Tijk_, const double *Zijk_) {
  const zd *epsi = (const zd *)epsi_, *Tijk = (const zd *)Tijk_, *Zijk = (const zd *)Zijk_;
  const long bs = 16, NN = No * No;
  const zd two = 2.0, three = 3.0, eabc = epsabc;
  zd energy = 0.0;
  for (long kk = 0; kk < No; kk += bs) {
    const long kend = kk + bs < No ? kk + bs : No;
    for (long jj = kk; jj < No; jj += bs) {
      const long jend = jj + bs < No ? jj + bs : No;
      for (long ii = jj; ii < No; ii += bs) {
        const long iend = ii + bs < No ? ii + bs : No;
        for (long k = kk; k < kend; k++)
          for (long j = jj > k ? jj : k; j < jend; j++) {
            const zd facjk = j == k ? 0.5 : 1.0;
            for (long i = ii > j ? ii : j; i < iend; i++) {
              const zd facij = i == j ? 0.5 : 1.0;
              const zd den = eabc - ((epsi[i] + epsi[j]) + epsi[k]);
              const zd U = Zijk[i + No * j + NN * k], V = Zijk[i + No * k + NN * j],
                       W = Zijk[j + No * i + NN * k], X = Zijk[j + No * k + NN * i],
                       Y = Zijk[k + No * i + NN * j], Z = Zijk[k + No * j + NN * i];
              const zd A = conj(Tijk[i + No * j + NN * k]), B = conj(Tijk[i + No * k + NN * j]),
                       C = conj(Tijk[j + No * i + NN * k]), D = conj(Tijk[j + No * k + NN * i]),
                       E = conj(Tijk[k + No * i + NN * j]), F = conj(Tijk[k + No * j + NN * i]);
              const zd UXY = U + (X + Y), VWZ = V + (W + Z);
              const zd ADE = A + (D + E), BCF = B + (C + F);
              const zd first = A * U + (B * V + (C * W + (D * X + (E * Y + F * Z))));
              const zd second = (UXY - two * VWZ) * ADE;
              const zd third = (VWZ - two * UXY) * BCF;
              const zd value = three * first + (second + third);
              energy += ((two * value) / den) * (facjk * facij);
            }
          }
      }
    }
  }
  return creal(energy);
}

/* get_energy_same<Complex>, Equations.cxx:182-238 */
double oracle_energy_same_z(double epsabc, long No, const double *epsi_,
                            const double *Tijk_, const double *Zijk_) {
  const zd *epsi = (const zd *)epsi_, *Tijk = (const zd *)Tijk_, *Zijk = (const zd *)Zijk_;
  const long bs = 16, NN = No * No;
  const zd two = 2.0, three = 3.0, eabc = epsabc;
  zd energy = 0.0;
  for (long kk = 0; kk < No; kk += bs) {
    const long kend = kk + bs < No ? kk + bs : No;
    for (long jj = kk; jj < No; jj += bs) {
      const long jend = jj + bs < No ? jj + bs : No;
      for (long ii = jj; ii < No; ii += bs) {
        const long iend = ii + bs < No ? ii + bs : No;
        for (long k = kk; k < kend; k++)
          for (long j = jj > k ? jj : k; j < jend; j++) {
            const zd facjk = j == k ? 0.5 : 1.0;
            for (long i = ii > j ? ii : j; i < iend; i++) {
              const zd facij = i == j ? 0.5 : 1.0;
              const zd den = eabc - ((epsi[i] + epsi[j]) + epsi[k]);
              const zd U = Zijk[i + No * j + NN * k], V = Zijk[j + No * k + NN * i],
                       W = Zijk[k + No * i + NN * j];
              const zd A = conj(Tijk[i + No * j + NN * k]), B = conj(Tijk[j + No * k + NN * i]),
                       C = conj(Tijk[k + No * i + NN * j]);
              const zd ABC = A + (B + C), UVW = U + (V + W);
              const zd value = three * ((A * U + B * V) + C * W) - ABC * UVW;
              energy += ((two * value) / den) * (facjk * facij);
            }
          }
      }
    }
  }
  return creal(energy);
}

/* one iteration of the main loop for F = Complex (Atrip.cxx:855-963); all
 * tensors interleaved complex.  Tijk_out / Zijk_out: 2 No^3 doubles each. */
double oracle_tuple_energy_z(long No, long Nv, const double *epsi, const double *epsa_,
                             const double *Tai, const double *Tabij_,
                             const double *Vabij_, const double *Vijka_,
                             const double *Vabci_, const double *Jijka_,
                             const double *Jabci_, long a, long b, long c,
                             double *Tijk_out, double *Zijk_out, double *ct) {
  const zd *epsa = (const zd *)epsa_, *Tabij = (const zd *)Tabij_, *Vabij = (const zd *)Vabij_,
           *Vijka = (const zd *)Vijka_, *Vabci = (const zd *)Vabci_, *Jijka = (const zd *)Jijka_,
           *Jabci = (const zd *)Jabci_;
  const long N3 = No * No * No, NvNo = Nv * No, NN = No * No;
  zd *buf = (zd *)malloc(sizeof(zd) * (6 * NvNo + 3 * N3 + 3 * NvNo * No + 6 * NN + 2 * N3));
  zd *VAB = buf, *VAC = VAB + NvNo, *VBC = VAC + NvNo, *VBA = VBC + NvNo, *VCA = VBA + NvNo,
     *VCB = VCA + NvNo;
  zd *HA = VCB + NvNo, *HB = HA + N3, *HC = HB + N3;
  zd *TA = HC + N3, *TB = TA + NvNo * No, *TC = TB + NvNo * No;
  zd *TAB = TC + NvNo * No, *TAC = TAB + NN, *TBC = TAC + NN;
  zd *VABij = TBC + NN, *VACij = VABij + NN, *VBCij = VACij + NN;
  zd *Tijk = VBCij + NN, *Zijk = Tijk + N3;
#define D(p) ((double *)(p))
  const double epsabc = creal(epsa[a] + epsa[b] + epsa[c]); /* Atrip.cxx:643-644 */
  const int same = (a == b) != (b == c);                    /* Atrip.cxx:640-642 */
  double e = 0.0;
  const int npass = (ct && Jijka && Jabci) ? 2 : 1;
  for (int pass = 0; pass < npass; pass++) {
    const zd *vp = pass ? Jabci : Vabci, *vh = pass ? Jijka : Vijka;
    slice_ABPH(No, Nv, vp, a, b, VAB);
    slice_ABPH(No, Nv, vp, a, c, VAC);
    slice_ABPH(No, Nv, vp, b, c, VBC);
    slice_ABPH(No, Nv, vp, b, a, VBA);
    slice_ABPH(No, Nv, vp, c, a, VCA);
    slice_ABPH(No, Nv, vp, c, b, VCB);
    slice_HHHA(No, vh, a, HA);
    slice_HHHA(No, vh, b, HB);
    slice_HHHA(No, vh, c, HC);
    if (pass == 0) {
      slice_TA(No, Nv, Tabij, a, TA);
      slice_TA(No, Nv, Tabij, b, TB);
      slice_TA(No, Nv, Tabij, c, TC);
      slice_ABHH(No, Nv, Tabij, a, b, TAB);
      slice_ABHH(No, Nv, Tabij, a, c, TAC);
      slice_ABHH(No, Nv, Tabij, b, c, TBC);
      slice_ABHH(No, Nv, Vabij, a, b, VABij);
      slice_ABHH(No, Nv, Vabij, a, c, VACij);
      slice_ABHH(No, Nv, Vabij, b, c, VBCij);
    }
    oracle_doubles_z(No, Nv, D(VAB), D(VAC), D(VBC), D(VBA), D(VCA), D(VCB), D(HA), D(HB), D(HC),
                     D(TA), D(TB), D(TC), D(TAB), D(TAC), D(TBC), D(Tijk));
    if (pass == 0) { /* Zijk is built once, from the V pass (Atrip.cxx:899-906) */
      memcpy(Zijk, Tijk, sizeof(zd) * N3);
      oracle_singles_z(No, Nv, a, b, c, Tai, D(VABij), D(VACij), D(VBCij), D(Zijk));
      if (Tijk_out) memcpy(Tijk_out, Tijk, sizeof(zd) * N3);
      if (Zijk_out) memcpy(Zijk_out, Zijk, sizeof(zd) * N3);
    }
    const double ep = same ? oracle_energy_same_z(epsabc, No, epsi, D(Tijk), D(Zijk))
                           : oracle_energy_distinct_z(epsabc, No, epsi, D(Tijk), D(Zijk));
    if (pass == 0) {
      e = ep;
      if (ct) *ct = ep; /* without J the reference evaluates the same energy twice */
    } else {
      *ct = ep; /* (cT): Tijk from the J pass, Zijk from the V pass (Atrip.cxx:928-963) */
    }
  }
#undef D
  free(buf);
  return e;
}

int oracle_run_z(long No, long Nv, const double *epsi, const double *epsa,
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
    e += oracle_tuple_energy_z(No, Nv, epsi, epsa, Tai, Tabij, Vabij, Vijka, Vabci, Jijka, Jabci,
                               a, b, c, NULL, NULL, &ct);
    ect += ct;
  }
  *energy = -e;
  *ct_energy = -ect;
  free(own);
  return 0;
}
