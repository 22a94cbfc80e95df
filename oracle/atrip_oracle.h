/* TEST INFRASTRUCTURE ONLY.  Plain-C CPU restatement of atrip's (T) hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  See atrip_oracle.c for the reference citations
 * and for how the restatement is pinned (oracle/_ref = the reference itself).
 */
#ifndef ATRIP_ORACLE_H
#define ATRIP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* synthetic inputs (DESIGN.md "Synthetic inputs") */
enum {
  ORACLE_EPS_I = 0,
  ORACLE_EPS_A = 1,
  ORACLE_TAI = 2,
  ORACLE_TABIJ = 3,
  ORACLE_VABIJ = 4,
  ORACLE_VIJKA = 5,
  ORACLE_VABCI = 6,
  ORACLE_JIJKA = 7,
  ORACLE_JABCI = 8
};
double oracle_synth(uint64_t seed, int tensor_id, uint64_t idx, double scale);
void oracle_fill(uint64_t seed, int tensor_id, double scale, uint64_t first,
                 uint64_t count, double *out);

/* slices (reference Unions.hpp:77-278) */
void oracle_slice_TA(long No, long Nv, const double *Tabij, long x, double *out);
void oracle_slice_HHHA(long No, long Nv, const double *Vijka, long x, double *out);
void oracle_slice_ABPH(long No, long Nv, const double *Vabci, long x, long y, double *out);
void oracle_slice_ABHH(long No, long Nv, const double *Vabij, long x, long y, double *out);

/* the 18 slices of one tuple from the generator alone (see atrip_oracle.c) */
void oracle_synth_tuple_slices(uint64_t seed, double scale, long No, long Nv, long a,
                               long b, long c, double **out);

/* per-tuple math (reference Equations.cxx) */
void oracle_doubles(long No, long Nv, const double *VAB, const double *VAC,
                    const double *VBC, const double *VBA, const double *VCA,
                    const double *VCB, const double *HA, const double *HB,
                    const double *HC, const double *TA, const double *TB,
                    const double *TC, const double *TAB, const double *TAC,
                    const double *TBC, double *Tijk);
void oracle_singles(long No, long Nv, long a, long b, long c, const double *Tph,
                    const double *VABij, const double *VACij,
                    const double *VBCij, double *Zijk);
double oracle_energy_distinct(double epsabc, long No, const double *epsi,
                              const double *Tijk, const double *Zijk);
double oracle_energy_same(double epsabc, long No, const double *epsi,
                          const double *Tijk, const double *Zijk);

/* one tuple from the full tensors; optionally returns Tijk / Zijk (No^3 each).
 * With Jijka/Jabci non-NULL also returns the (cT) tuple energy in *ct. */
double oracle_tuple_energy(long No, long Nv, const double *epsi, const double *epsa,
                           const double *Tai, const double *Tabij,
                           const double *Vabij, const double *Vijka,
                           const double *Vabci, const double *Jijka,
                           const double *Jabci, long a, long b, long c,
                           double *Tijk_out, double *Zijk_out, double *ct);

/* tuples (reference Tuples.cxx) */
long oracle_n_tuples(long Nv);
long oracle_all_tuples(long Nv, uint64_t *out, long cap);
long oracle_group_and_sort(long n_nodes, long node_id, long Nv, uint64_t *out, long cap);

/* slice ownership (reference RankMap.cxx:35-85 with one rank per node) */
long oracle_owner_single(long x, long n_ranks);
long oracle_owner_pair(long x, long y, long Nv, long n_ranks);

/* whole run: -sum over the given tuples (all tuples when tuples == NULL) */
int oracle_run(long No, long Nv, const double *epsi, const double *epsa,
               const double *Tai, const double *Tabij, const double *Vabij,
               const double *Vijka, const double *Vabci, const double *Jijka,
               const double *Jabci, const uint64_t *tuples, long n_tuples,
               double *energy, double *ct_energy);

/* ---- F = std::complex<double> (atrip_oracle_z.c); arrays are interleaved (re, im) */
void oracle_fill_z(uint64_t seed, int tensor_id, double scale, uint64_t first,
                   uint64_t count, double *out);
void oracle_doubles_z(long No, long Nv, const double *VAB, const double *VAC,
                      const double *VBC, const double *VBA, const double *VCA,
                      const double *VCB, const double *HA, const double *HB,
                      const double *HC, const double *TA, const double *TB,
                      const double *TC, const double *TAB, const double *TAC,
                      const double *TBC, double *Tijk);
void oracle_singles_z(long No, long Nv, long a, long b, long c, const double *Tph,
                      const double *VABij, const double *VACij,
                      const double *VBCij, double *Zijk);
double oracle_energy_distinct_z(double epsabc, long No, const double *epsi,
                                const double *Tijk, const double *Zijk);
double oracle_energy_same_z(double epsabc, long No, const double *epsi,
                            const double *Tijk, const double *Zijk);
double oracle_tuple_energy_z(long No, long Nv, const double *epsi, const double *epsa,
                             const double *Tai, const double *Tabij,
                             const double *Vabij, const double *Vijka,
                             const double *Vabci, const double *Jijka,
                             const double *Jabci, long a, long b, long c,
                             double *Tijk_out, double *Zijk_out, double *ct);
int oracle_run_z(long No, long Nv, const double *epsi, const double *epsa,
                 const double *Tai, const double *Tabij, const double *Vabij,
                 const double *Vijka, const double *Vabci, const double *Jijka,
                 const double *Jabci, const uint64_t *tuples, long n_tuples,
                 double *energy, double *ct_energy);

#ifdef __cplusplus
}
#endif
#endif
