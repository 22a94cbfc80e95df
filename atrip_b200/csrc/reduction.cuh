// Kernel 2 of the (T) hot path: singles + energy denominators + tuple energy, fused.
//
// Replaces, per tuple, xcopy Zijk = Tijk (reference Atrip.cxx:899-906), singles_contribution
// (Equations.cxx:387-426) and get_energy_distinct / get_energy_same (Equations.cxx:101-238),
// which the reference's GPU path runs as <<<1,1>>> kernels plus a cuMemAlloc and a blocking
// 8-byte DtoH per tuple (Atrip.cxx:630-676).  Neither Tijk nor Zijk is materialised: one CTA
// per (tuple, orbit split) walks the orbits {I >= J >= K} of 8x8x8 tiles of the occupied cube,
// rebuilds
//     Tijk[x,y,z] = C_k[x,y,z] + C_j[x,y,z] + C_i[x,y,z]
// (kernel 1 stores its three class cubes in Tijk's own index order) for the six permuted tiles of
// the orbit in shared memory -- every global load of an orbit is issued before the first use, so
// a thread keeps up to 36 loads in flight --, forms Zijk on the fly from Tai rows and the three Vabij
// blocks, evaluates the reference's triangular sum k <= j <= i with its weights literally
// (needed for the a==b / b==c tuples on unsymmetric data), and reduces
// with warp shuffles to one double per tuple.  The batch sum is a second, single-CTA kernel in a
// fixed order, so results are run-to-run deterministic and there is no per-tuple DtoH.
#pragma once
#include <utility>

#include "common.cuh"
#include "contraction.cuh"
#include "schedule.hpp"

namespace ab {

constexpr int RT = 8;                 // tile edge
constexpr int RTILE = 8 * 72;         // doubles per tile in shared memory: z planes padded to 72
constexpr int REDUCE_THREADS = 128;   // 4 warps: one CTA fits beside a contraction CTA on an SM

// Shared-memory position of tile element (x,y,z): the x index is XOR-swizzled with (y & 6) ^ z and
// the z planes are padded to 72 doubles.  A half-warp of the compute phase (8 values of i, 2 of
// j, one k) then hits 16 different 8-byte banks for ALL six index permutations [i,j,k] ... [k,j,i]
// (ncu r01b: the former [x + 9 y + 81 z] layout replayed every LDS 2.2 times).
__device__ __forceinline__ int tile_pos(int x, int y, int z) { return (x ^ ((y & 6) ^ z)) + 8 * y + 72 * z; }

struct ReduceParams {
  int No, Nv;
  int ntuples;
  const TupleRec *recs;  // the batch: tuple + store slots (vij: Vabij blocks of (b,c) (a,c) (a,b))
  const double *R;    // class cubes giving Tijk           [ntuples][3][cube_stride], 8x8x8-blocked
  size_t cube_stride; // cube_blocked_elems(No)
  const double *RZ;   // class cubes giving the Tijk inside Zijk (== R except in the cT pass)
  const double *eps_i, *eps_a, *Tai;
  const double *VIJ;  // owned Vabij pair blocks [slot][No^2]
  const double *VIJc; // fetch cache, addressed by slot - ownedV
  int ownedV;
  double *e_tuple;    // [ntuples * nsplit] partial energies
  int nsplit;         // CTAs per tuple: CTA (t, s) takes the orbits o with o % nsplit == s
  int reverse;        // experimental kernels only: walk the batch's tuples last-to-first (the cubes the
                      // contraction wrote last are the ones still resident in L2)
};

__host__ __device__ inline size_t reduce_smem_bytes(int No, bool ct) {
  return sizeof(double) * ((size_t)(ct ? 12 : 6) * RTILE + 18 * 64 + 4 * (size_t)No + 32);
}

// compile-time loop over the six permuted tiles: body(std::integral_constant<int, p>) with
// p = 0 (I,J,K) 1 (I,K,J) 2 (J,I,K) 3 (J,K,I) 4 (K,I,J) 5 (K,J,I).  Everything that depends on p
// is a constant in each instantiation, so the staging arrays stay in registers (with a run-time p
// nvcc indexed the tables through selects and demoted every array to local memory).
template <int P>
struct TilePerm {
  static constexpr int X = P >> 1;
  static constexpr int Y = (P == 0 || P == 5) ? 1 : ((P == 1 || P == 3) ? 2 : 0);
  static constexpr int Z = 3 - X - Y;
};
__device__ __forceinline__ int pick3(int which, int I, int J, int K) { return which == 0 ? I : (which == 1 ? J : K); }
template <typename F, int... Ps>
__device__ __forceinline__ void for_tiles_impl(F &&f, std::integer_sequence<int, Ps...>) {
  (f(std::integral_constant<int, Ps>{}), ...);
}
template <typename F>
__device__ __forceinline__ void for_tiles(F &&f) {
  for_tiles_impl(f, std::make_integer_sequence<int, 6>{});
}

#ifndef REDUCE_MINBLOCKS
#define REDUCE_MINBLOCKS 4
#endif
template <bool CT>
__global__ void __launch_bounds__(REDUCE_THREADS, REDUCE_MINBLOCKS)
reduce_kernel(const ReduceParams P) {
  extern __shared__ double sm[];
  double *Wt = sm;                               // [6][RTILE]
  double *Zt = CT ? sm + 6 * RTILE : sm;         // [6][RTILE] (aliases Wt when !CT)
  double *Vb = sm + (CT ? 12 : 6) * RTILE;       // [3][6][64]
  double *sEps = Vb + 18 * 64;                   // [No]
  double *sTa = sEps + P.No, *sTb = sTa + P.No, *sTc = sTb + P.No;
  double *sRed = sTc + P.No;                     // [32]

  const int tup = blockIdx.x;
  const TupleRec rec = P.recs[tup];
  const int tid = threadIdx.x;
  const int split = blockIdx.y;
  if (rec.fake) {  // FAKE_TUPLE contributes nothing (Atrip.cxx:629)
    if (tid == 0) P.e_tuple[(size_t)tup * P.nsplit + split] = 0.0;
    return;
  }
  const int a = rec.a, b = rec.b, c = rec.c;
  const int No = P.No, Nv = P.Nv;
  const size_t NoNo = (size_t)No * No, cube = P.cube_stride;
  // the three class cubes of kernel 1, all at Tijk's (i,j,k), stored as contiguous 8x8x8 tiles
  const double *Ck = P.R + (size_t)tup * 3 * cube, *Cj = Ck + cube, *Ci = Cj + cube;
  const double *Zk = P.RZ + (size_t)tup * 3 * cube, *Zj = Zk + cube, *Zi = Zj + cube;
  const double *Vmat[3];  // VBCij, VACij, VABij
#pragma unroll
  for (int q = 0; q < 3; q++)
    Vmat[q] = rec.vij[q] >= P.ownedV ? P.VIJc + (size_t)(rec.vij[q] - P.ownedV) * NoNo : P.VIJ + (size_t)rec.vij[q] * NoNo;
  for (int i = tid; i < No; i += REDUCE_THREADS) {
    sEps[i] = P.eps_i[i];
    sTa[i] = P.Tai[a + (size_t)i * Nv];
    sTb[i] = P.Tai[b + (size_t)i * Nv];
    sTc[i] = P.Tai[c + (size_t)i * Nv];
  }
  const double epsabc = P.eps_a[a] + P.eps_a[b] + P.eps_a[c];  // Atrip.cxx:643-646
  const bool same = (a == b) != (b == c);                     // Atrip.cxx:640-650

  const int nb = (No + RT - 1) / RT;
  double esum = 0.0;
  // element e = tid + 128 q of a tile is (x,y,z) = (l0, l1, l2 + 2 q); the same mapping gives
  // the 4 points (il, jl, kl + 2 q) a thread evaluates
  const int l0 = tid & 7, l1 = (tid >> 3) & 7, l2 = tid >> 6;  // l2 in 0..1

  __syncthreads();  // publishes sEps, sTa, sTb, sTc
  int orbit = -1;
  for (int I = 0; I < nb; I++)
    for (int J = 0; J <= I; J++)
      for (int K = 0; K <= J; K++) {
        if (++orbit % P.nsplit != split) continue;
        // coincident block coordinates give identical tiles: build each distinct one once
        // (tile p -> its first equal tile c_p)
        const bool eIJ = (I == J), eJK = (J == K);
        const int c1 = eJK ? 0 : 1, c2 = eIJ ? 0 : 2, c3 = (eIJ && eJK) ? 0 : (eIJ ? 1 : 3),
                  c4 = (eIJ && eJK) ? 0 : (eJK ? 2 : 4), c5 = (eIJ && eJK) ? 0 : (eIJ ? 4 : (eJK ? 3 : 5));
        const bool d1 = c1 == 1, d2 = c2 == 2, d3 = c3 == 3, d4 = c4 == 4, d5 = c5 == 5;
        auto distinct = [&](int p) { return p == 0 || (p == 1 ? d1 : (p == 2 ? d2 : (p == 3 ? d3 : (p == 4 ? d4 : d5)))); };
        // ---- issue every global load of the orbit before touching shared memory: up to
        //      6 tiles x 4 elements x 3 class cubes per thread in flight (predicated, no branches);
        //      a tile is 512 contiguous doubles, so these are full-line coalesced loads
        double lk[6][4], lj[6][4], li[6][4];
        for_tiles([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          using T = TilePerm<p>;
          const bool ok = distinct(p);
          const size_t tb =
              (((size_t)pick3(T::Z, I, J, K) * nb + pick3(T::Y, I, J, K)) * nb + pick3(T::X, I, J, K)) * 512 + tid;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            lk[p][q] = ok ? Ck[tb + 128 * q] : 0.0;
            lj[p][q] = ok ? Cj[tb + 128 * q] : 0.0;
            li[p][q] = ok ? Ci[tb + 128 * q] : 0.0;
          }
        });
        // Vabij blocks: Vb[mat][pair(X,Y)][xl + 8 yl] = Vmat[x + y No]; pair blocks
        // 0 (I,J) 1 (I,K) 2 (J,I) 3 (J,K) 4 (K,I) 5 (K,J)
        double lv[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
          const int e = tid + REDUCE_THREADS * q;  // < 18 * 64 = 9 * 128
          const int mat = e / 384, r = e - mat * 384, pr = r >> 6, xl = r & 7, yl = (r >> 3) & 7;
          const int X = pr >> 1, Y = (pr & 1) ? (X == 2 ? 1 : 2) : (X == 0 ? 1 : 0);
          const int x = pick3(X, I, J, K) * RT + xl, y = pick3(Y, I, J, K) * RT + yl;
          const double *vm = mat == 0 ? Vmat[0] : (mat == 1 ? Vmat[1] : Vmat[2]);
          lv[q] = (x < No && y < No) ? vm[x + (size_t)y * No] : 0.0;
        }
        __syncthreads();  // previous orbit fully consumed
        for_tiles([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          if (distinct(p)) {
#pragma unroll
            for (int q = 0; q < 4; q++) Wt[p * RTILE + tile_pos(l0, l1, l2 + 2 * q)] = (lk[p][q] + lj[p][q]) + li[p][q];
          }
        });
#pragma unroll
        for (int q = 0; q < 9; q++) Vb[tid + REDUCE_THREADS * q] = lv[q];
        if (CT) {  // (cT): Zijk comes from the V-pass cubes, Tijk (above) from the J pass
          for_tiles([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            using T = TilePerm<p>;
            if (distinct(p)) {
              const size_t tb =
                  (((size_t)pick3(T::Z, I, J, K) * nb + pick3(T::Y, I, J, K)) * nb + pick3(T::X, I, J, K)) * 512 + tid;
#pragma unroll
              for (int q = 0; q < 4; q++)
                Zt[p * RTILE + tile_pos(l0, l1, l2 + 2 * q)] = (Zk[tb + 128 * q] + Zj[tb + 128 * q]) + Zi[tb + 128 * q];
            }
          });
        }
        __syncthreads();
        // ---- energy of the (i in I, j in J, k in K) points with k <= j <= i
        const int il = l0, jl = l1;
        const int i = I * RT + il, j = J * RT + jl;
        if (i < No && j <= i) {
          const double *Vbc = Vb, *Vac = Vb + 384, *Vab = Vb + 768;
          // Vabij pair blocks: 0 (I,J) 1 (I,K) 2 (J,I) 3 (J,K) 4 (K,I) 5 (K,J)
          const int pij = 0 * 64 + il + 8 * jl, pji = 2 * 64 + jl + 8 * il;
          const double tai = sTa[i], taj = sTa[j], tbi = sTb[i], tbj = sTb[j], tci = sTc[i], tcj = sTc[j];
          const double eij = sEps[i] + sEps[j];
          const double facij = (i == j) ? 0.5 : 1.0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int kl = l2 + 2 * q, k = K * RT + kl;
            if (k <= j) {
              // tiles: 0 (I,J,K) 1 (I,K,J) 2 (J,I,K) 3 (J,K,I) 4 (K,I,J) 5 (K,J,I)
              const int o0 = tile_pos(il, jl, kl), o1 = RTILE * c1 + tile_pos(il, kl, jl);
              const int o2 = RTILE * c2 + tile_pos(jl, il, kl), o3 = RTILE * c3 + tile_pos(jl, kl, il);
              const int o4 = RTILE * c4 + tile_pos(kl, il, jl), o5 = RTILE * c5 + tile_pos(kl, jl, il);
              const double A = Wt[o0], B = Wt[o1], C = Wt[o2], D = Wt[o3], E = Wt[o4], F = Wt[o5];
              const int pik = 1 * 64 + il + 8 * kl, pjk = 3 * 64 + jl + 8 * kl;
              const int pki = 4 * 64 + kl + 8 * il, pkj = 5 * 64 + kl + 8 * jl;
              const double tak = sTa[k], tbk = sTb[k], tck = sTc[k];
              // Z[x,y,z] = T[x,y,z] + Tai[a,x] Vbc[y,z] + Tai[b,y] Vac[x,z] + Tai[c,z] Vab[x,y]
              // (Equations.cxx:420-422, three separate += in this order)
              double U = Zt[o0], V = Zt[o1], W = Zt[o2], X = Zt[o3], Y = Zt[o4], Z = Zt[o5];
              U = ((U + tai * Vbc[pjk]) + tbj * Vac[pik]) + tck * Vab[pij];  // Z[i,j,k]
              V = ((V + tai * Vbc[pkj]) + tbk * Vac[pij]) + tcj * Vab[pik];  // Z[i,k,j]
              W = ((W + taj * Vbc[pik]) + tbi * Vac[pjk]) + tck * Vab[pji];  // Z[j,i,k]
              X = ((X + taj * Vbc[pki]) + tbk * Vac[pji]) + tci * Vab[pjk];  // Z[j,k,i]
              Y = ((Y + tak * Vbc[pij]) + tbi * Vac[pkj]) + tcj * Vab[pki];  // Z[k,i,j]
              Z = ((Z + tak * Vbc[pji]) + tbj * Vac[pki]) + tci * Vab[pkj];  // Z[k,j,i]
              const double facjk = (j == k) ? 0.5 : 1.0;
              const double den = epsabc - (eij + sEps[k]);
              double value;
              if (!same) {  // get_energy_distinct, Equations.cxx:129-166
                const double UXY = U + (X + Y), VWZ = V + (W + Z);
                const double ADE = A + (D + E), BCF = B + (C + F);
                const double first = A * U + (B * V + (C * W + (D * X + (E * Y + F * Z))));
                const double second = (UXY - 2.0 * VWZ) * ADE;
                const double third = (VWZ - 2.0 * UXY) * BCF;
                value = 3.0 * first + (second + third);
              } else {  // get_energy_same, Equations.cxx:209-226: cyclic permutations only
                const double ABC = A + (D + E), UVW = U + (X + Y);
                value = 3.0 * ((A * U + D * X) + E * Y) - ABC * UVW;
              }
              esum += ((2.0 * value) / den) * (facjk * facij);
            }
          }
        }
      }

  // block reduction in a fixed order
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) esum += __shfl_down_sync(0xffffffffu, esum, off);
  __syncthreads();
  if ((tid & 31) == 0) sRed[tid >> 5] = esum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < REDUCE_THREADS / 32; w++) s += sRed[w];
    P.e_tuple[(size_t)tup * P.nsplit + split] = s;
  }
}

// total[0] += sum_t e_tuple[t] in a fixed order (one CTA); keeps the energy on the device
__global__ void __launch_bounds__(256) accumulate_kernel(const double *e_tuple, int n, double *total) {
  __shared__ double s[256];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) v += e_tuple[i];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] += s[0];
}

// debug / parity only: materialise the reference's Tijk and Zijk for one tuple of the batch
__global__ void cubes_kernel(const ReduceParams P, int tup, double *Tijk, double *Zijk) {
  const int No = P.No, Nv = P.Nv;
  const size_t NoNo = (size_t)No * No, cube = NoNo * No;
  const TupleRec rec = P.recs[tup];
  const int3 abc = make_int3(rec.a, rec.b, rec.c);
  const double *Ck = P.R + (size_t)tup * 3 * P.cube_stride, *Cj = Ck + P.cube_stride, *Ci = Cj + P.cube_stride;
  const double *Vm[3];
  for (int q = 0; q < 3; q++)
    Vm[q] = rec.vij[q] >= P.ownedV ? P.VIJc + (size_t)(rec.vij[q] - P.ownedV) * NoNo : P.VIJ + (size_t)rec.vij[q] * NoNo;
  const double *Vbc = Vm[0], *Vac = Vm[1], *Vab = Vm[2];
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < cube; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % No), j = (int)((e / No) % No), k = (int)(e / NoNo);
    const size_t o = cube_offset(No, i, j, k);  // the class cubes share Tijk's (i,j,k)
    const double w = (Ck[o] + Cj[o]) + Ci[o];
    if (Tijk) Tijk[e] = w;
    if (Zijk)
      Zijk[e] = ((w + P.Tai[abc.x + (size_t)i * Nv] * Vbc[j + (size_t)k * No]) +
                 P.Tai[abc.y + (size_t)j * Nv] * Vac[i + (size_t)k * No]) +
                P.Tai[abc.z + (size_t)k * Nv] * Vab[i + (size_t)j * No];
  }
}

}  // namespace ab
