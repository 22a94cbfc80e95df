// Kernel 2, bulk-copy variant: the default (T)-pass reduction of the real field (ATRIP_B200_REDUCE=sync
// selects reduce_kernel of reduction.cuh instead; the (cT) pass and the complex field use that one).
//
// Same mathematics, orbit walk, tile layout and per-point operation order as reduce_kernel
// (reduction.cuh) -- only the way the class-cube tiles reach shared memory differs.  ncu of
// reduce_kernel (profiles/r01_stall_analysis_v2.txt) shows 36 % of the stall samples on the local
// stores that spill its 72 register-staged loads per thread, and no load in flight during the
// energy phase of a CTA.  Here the tiles of the NEXT orbit travel as 4 KB cp.async.bulk copies
// (one elected thread, mbarrier completion) into a raw staging area while the CTA evaluates the
// current orbit from the summed tiles:
//     wait(staging of orbit n) -> sum the three classes into the swizzled W tiles -> barrier
//     -> issue the bulk copies of orbit n+1 into the (now free) staging area -> energy of orbit n.
// No register staging, no spills, loads always in flight; 110 KB of shared memory, 2 CTAs of 256 threads per SM.
// Measured (round 2): 236 us per c2 launch (699 tuples, 1.10 GB -> 4.7 TB/s, profiles/r02i_reduce_async_c2_ncu.txt)
// against 349 us of reduce_kernel; whole runs +2.5 % (c2), +1.5 % (c3), +1.7 % (c4 shapes).
#pragma once
#include "reduction.cuh"

namespace ab {

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int RA_STAGE_DOUBLES = 18 * 512;  // 6 tiles x 3 classes x 4 KB
constexpr int RA_THREADS = 256;              // 8 warps: 2 points of a tile per thread (128 and 256 both work; 256
                                             // hides the shared-memory latency better: 4.2 -> 4.7 TB/s on c2)

__host__ __device__ inline size_t reduce_async_smem_bytes(int No) {
  return sizeof(double) * ((size_t)RA_STAGE_DOUBLES + 6 * RTILE + 18 * 64 + 4 * (size_t)No + 32) + 16;
}

// the orbits {I >= J >= K} of 8x8x8 tiles assigned to one CTA: o = split, split + nsplit, ...
struct OrbitWalk {
  int I = 0, J = 0, K = 0, nb, nsplit;
  bool valid;
  __device__ OrbitWalk(int nb_, int nsplit_, int split) : nb(nb_), nsplit(nsplit_), valid(nb_ > 0) { skip(split); }
  __device__ void step() {
    if (++K > J) {
      K = 0;
      if (++J > I) {
        J = 0;
        if (++I >= nb) valid = false;
      }
    }
  }
  __device__ void skip(int n) {
    for (int s = 0; s < n && valid; s++) step();
  }
  __device__ void next() { skip(nsplit); }
};

__global__ void __launch_bounds__(RA_THREADS, 2) reduce_async_kernel(const ReduceParams P) {
  constexpr int NQ = 512 / RA_THREADS;  // tile elements (and energy points) per thread
  extern __shared__ __align__(16) double sm_async[];
  double *St = sm_async;                              // [6 tiles][3 classes][512] raw class tiles of one orbit
  double *Wt = St + RA_STAGE_DOUBLES;           // [6][RTILE] summed, swizzled
  double *Vb = Wt + 6 * RTILE;                  // [3][6][64]
  double *sEps = Vb + 18 * 64;                  // [No]
  double *sTa = sEps + P.No, *sTb = sTa + P.No, *sTc = sTb + P.No;
  double *sRed = sTc + P.No;                    // [32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(sRed + 32);

  const int tup = P.reverse ? P.ntuples - 1 - (int)blockIdx.x : (int)blockIdx.x;
  const TupleRec rec = P.recs[tup];
  const int tid = threadIdx.x;
  const int split = blockIdx.y;
  if (rec.fake) {
    if (tid == 0) P.e_tuple[(size_t)tup * P.nsplit + split] = 0.0;
    return;
  }
  const int a = rec.a, b = rec.b, c = rec.c;
  const int No = P.No, Nv = P.Nv;
  const size_t NoNo = (size_t)No * No, cube = P.cube_stride;
  const double *Ck = P.R + (size_t)tup * 3 * cube;  // classes at Ck, Ck + cube, Ck + 2 cube
  const double *Vmat[3];
#pragma unroll
  for (int q = 0; q < 3; q++)
    Vmat[q] = rec.vij[q] >= P.ownedV ? P.VIJc + (size_t)(rec.vij[q] - P.ownedV) * NoNo : P.VIJ + (size_t)rec.vij[q] * NoNo;
  for (int i = tid; i < No; i += RA_THREADS) {
    sEps[i] = P.eps_i[i];
    sTa[i] = P.Tai[a + (size_t)i * Nv];
    sTb[i] = P.Tai[b + (size_t)i * Nv];
    sTc[i] = P.Tai[c + (size_t)i * Nv];
  }
  const double epsabc = P.eps_a[a] + P.eps_a[b] + P.eps_a[c];
  const bool same = (a == b) != (b == c);
  const int nb = (No + RT - 1) / RT;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();

  // distinct tiles of an orbit (tile p -> first equal tile), as in reduce_kernel
  auto canon = [](int I, int J, int K, int cm[6]) {
    const bool eIJ = (I == J), eJK = (J == K);
    cm[0] = 0;
    cm[1] = eJK ? 0 : 1;
    cm[2] = eIJ ? 0 : 2;
    cm[3] = (eIJ && eJK) ? 0 : (eIJ ? 1 : 3);
    cm[4] = (eIJ && eJK) ? 0 : (eJK ? 2 : 4);
    cm[5] = (eIJ && eJK) ? 0 : (eIJ ? 4 : (eJK ? 3 : 5));
  };
  // elected thread: bulk copies of the distinct tiles of orbit (I,J,K), three classes each
  auto issue = [&](int I, int J, int K) {
    int cm[6];
    canon(I, J, K, cm);
    const int X[6] = {I, I, J, J, K, K}, Y[6] = {J, K, I, K, I, J}, Z[6] = {K, J, K, I, J, I};
    int ntiles = 0;
#pragma unroll
    for (int p = 0; p < 6; p++) ntiles += (cm[p] == p);
    mbar_expect_tx(bar, (uint32_t)ntiles * 3u * 4096u);
#pragma unroll
    for (int p = 0; p < 6; p++) {
      if (cm[p] != p) continue;
      const size_t tb = (((size_t)Z[p] * nb + Y[p]) * nb + X[p]) * 512;
#pragma unroll
      for (int cls = 0; cls < 3; cls++) bulk_load_1d(St + (p * 3 + cls) * 512, Ck + cls * cube + tb, 4096u, bar);
    }
  };
  // Vabij pair blocks of an orbit into registers (stored to Vb when the orbit becomes current)
  constexpr int NV = (18 * 64 + RA_THREADS - 1) / RA_THREADS;  // Vabij block elements per thread
  auto load_v = [&](int I, int J, int K, double lv[NV]) {
#pragma unroll
    for (int q = 0; q < NV; q++) {
      const int e = min(tid + RA_THREADS * q, 18 * 64 - 1);
      const int mat = e / 384, r = e - mat * 384, pr = r >> 6, xl = r & 7, yl = (r >> 3) & 7;
      const int Xs = pr >> 1, Ys = (pr & 1) ? (Xs == 2 ? 1 : 2) : (Xs == 0 ? 1 : 0);
      const int x = pick3(Xs, I, J, K) * RT + xl, y = pick3(Ys, I, J, K) * RT + yl;
      const double *vm = mat == 0 ? Vmat[0] : (mat == 1 ? Vmat[1] : Vmat[2]);
      lv[q] = (x < No && y < No) ? vm[x + (size_t)y * No] : 0.0;
    }
  };

  OrbitWalk cur(nb, P.nsplit, split);
  double lv[NV];
  if (cur.valid) {
    if (tid == 0) issue(cur.I, cur.J, cur.K);
    load_v(cur.I, cur.J, cur.K, lv);
  }
  double esum = 0.0;
  const int l0 = tid & 7, l1 = (tid >> 3) & 7, l2 = tid >> 6;  // l2 < RA_THREADS / 64; points (l0, l1, l2 + (8 / NQ) q)
  constexpr int ZS = 8 / NQ;
  uint32_t phase = 0;
  while (cur.valid) {
    const int I = cur.I, J = cur.J, K = cur.K;
    int cm[6];
    canon(I, J, K, cm);
    const int c1 = cm[1], c2 = cm[2], c3 = cm[3], c4 = cm[4], c5 = cm[5];
    // ---- staging of this orbit has landed: sum the classes into the swizzled tiles
    mbar_wait(bar, phase);
    phase ^= 1;
#pragma unroll
    for (int p = 0; p < 6; p++) {
      if (cm[p] != p) continue;
      const double *s = St + p * 3 * 512 + tid;
#pragma unroll
      for (int q = 0; q < NQ; q++)
        Wt[p * RTILE + tile_pos(l0, l1, l2 + ZS * q)] =
            (s[RA_THREADS * q] + s[512 + RA_THREADS * q]) + s[1024 + RA_THREADS * q];
    }
#pragma unroll
    for (int q = 0; q < NV; q++)
      if (tid + RA_THREADS * q < 18 * 64) Vb[tid + RA_THREADS * q] = lv[q];
    __syncthreads();  // tiles and Vb published; staging free again
    // ---- next orbit: start its copies now, they fly during the energy phase below
    cur.next();
    if (cur.valid) {
      if (tid == 0) {
        fence_proxy_async();  // generic-proxy reads of the staging area above before async-proxy writes
        issue(cur.I, cur.J, cur.K);
      }
      load_v(cur.I, cur.J, cur.K, lv);
    }
    // ---- energy of the (i in I, j in J, k in K) points with k <= j <= i  (as reduce_kernel)
    const int il = l0, jl = l1;
    const int i = I * RT + il, j = J * RT + jl;
    if (i < No && j <= i) {
      const double *Vbc = Vb, *Vac = Vb + 384, *Vab = Vb + 768;
      const int pij = 0 * 64 + il + 8 * jl, pji = 2 * 64 + jl + 8 * il;
      const double tai = sTa[i], taj = sTa[j], tbi = sTb[i], tbj = sTb[j], tci = sTc[i], tcj = sTc[j];
      const double eij = sEps[i] + sEps[j];
      const double facij = (i == j) ? 0.5 : 1.0;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int kl = l2 + ZS * q, k = K * RT + kl;
        if (k <= j) {
          const int o0 = tile_pos(il, jl, kl), o1 = RTILE * c1 + tile_pos(il, kl, jl);
          const int o2 = RTILE * c2 + tile_pos(jl, il, kl), o3 = RTILE * c3 + tile_pos(jl, kl, il);
          const int o4 = RTILE * c4 + tile_pos(kl, il, jl), o5 = RTILE * c5 + tile_pos(kl, jl, il);
          const double A = Wt[o0], B = Wt[o1], C = Wt[o2], D = Wt[o3], E = Wt[o4], F = Wt[o5];
          const int pik = 1 * 64 + il + 8 * kl, pjk = 3 * 64 + jl + 8 * kl;
          const int pki = 4 * 64 + kl + 8 * il, pkj = 5 * 64 + kl + 8 * jl;
          const double tak = sTa[k], tbk = sTb[k], tck = sTc[k];
          double U = A, V = B, W = C, X = D, Y = E, Z = F;
          U = ((U + tai * Vbc[pjk]) + tbj * Vac[pik]) + tck * Vab[pij];  // Z[i,j,k]
          V = ((V + tai * Vbc[pkj]) + tbk * Vac[pij]) + tcj * Vab[pik];  // Z[i,k,j]
          W = ((W + taj * Vbc[pik]) + tbi * Vac[pjk]) + tck * Vab[pji];  // Z[j,i,k]
          X = ((X + taj * Vbc[pki]) + tbk * Vac[pji]) + tci * Vab[pjk];  // Z[j,k,i]
          Y = ((Y + tak * Vbc[pij]) + tbi * Vac[pkj]) + tcj * Vab[pki];  // Z[k,i,j]
          Z = ((Z + tak * Vbc[pji]) + tbj * Vac[pki]) + tci * Vab[pkj];  // Z[k,j,i]
          const double facjk = (j == k) ? 0.5 : 1.0;
          const double den = epsabc - (eij + sEps[k]);
          double value;
          if (!same) {
            const double UXY = U + (X + Y), VWZ = V + (W + Z);
            const double ADE = A + (D + E), BCF = B + (C + F);
            const double first = A * U + (B * V + (C * W + (D * X + (E * Y + F * Z))));
            const double second = (UXY - 2.0 * VWZ) * ADE;
            const double third = (VWZ - 2.0 * UXY) * BCF;
            value = 3.0 * first + (second + third);
          } else {
            const double ABC = A + (D + E), UVW = U + (X + Y);
            value = 3.0 * ((A * U + D * X) + E * Y) - ABC * UVW;
          }
          esum += ((2.0 * value) / den) * (facjk * facij);
        }
      }
    }
    __syncthreads();  // everybody is done with Wt / Vb before the next orbit rewrites them
  }

#pragma unroll
  for (int off = 16; off > 0; off >>= 1) esum += __shfl_down_sync(0xffffffffu, esum, off);
  if ((tid & 31) == 0) sRed[tid >> 5] = esum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < RA_THREADS / 32; w++) s += sRed[w];
    P.e_tuple[(size_t)tup * P.nsplit + split] = s;
  }
}

}  // namespace ab
