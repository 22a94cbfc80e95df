// Kernel 2 for the complex field (F = std::complex<double>): singles + denominators + energy.
//
// Same replacement and the same orbit/tile walk as reduction.cuh (reference Atrip.cxx:899-906,
// Equations.cxx:387-426, 101-238), instantiated for complex numbers: the contraction kernel ran
// twice per tuple (AX variants 0 and 1, stores.cuh "complex field") and left SIX class cubes,
// Re C_k, Re C_j, Re C_i, Im C_k, Im C_j, Im C_i, all at Tijk's (i,j,k).  Here
//     Tijk = sum of the three classes (re and im), Zijk = Tijk + Tai x Vabij (plain products,
//     Equations.cxx:420-422), and the energy sums use conj(Tijk) (Equations.cxx:135-146,
//     207-212), a complex denominator epsabc - (eps_i + eps_j + eps_k) with real epsabc
//     (Atrip.cxx:643-646) and keep the real part of the sum (:176-178, :234-236).
// The per-point arithmetic lives in __host__ __device__ functions so that the CPU suite can check
// it against the oracle without a GPU (atrip_b200_host_energy_z).
#pragma once
#include "reduction.cuh"

namespace ab {

struct cz {
  double re, im;
};
__host__ __device__ __forceinline__ cz operator+(cz a, cz b) { return cz{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cz operator-(cz a, cz b) { return cz{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cz operator*(cz a, cz b) {
  return cz{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ cz operator*(double s, cz a) { return cz{s * a.re, s * a.im}; }
__host__ __device__ __forceinline__ cz conj(cz a) { return cz{a.re, -a.im}; }
__host__ __device__ __forceinline__ cz operator/(cz a, cz b) {
  const double d = b.re * b.re + b.im * b.im;
  return cz{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// One point (i,j,k), k <= j <= i, of the tuple energy.  T[] / Z[] hold Tijk / Zijk at the six
// permutations 0 [i,j,k] 1 [i,k,j] 2 [j,i,k] 3 [j,k,i] 4 [k,i,j] 5 [k,j,i] (T not yet conjugated).
// Returns the complex summand ((2 value) / den) * (facjk facij); the caller keeps the real part
// of the sum.
__host__ __device__ __forceinline__ cz point_energy_z(bool same, const cz T[6], const cz Zp[6], cz den, double fac) {
  cz value;
  if (!same) {  // get_energy_distinct<Complex>, Equations.cxx:129-166
    const cz A = conj(T[0]), B = conj(T[1]), C = conj(T[2]), D = conj(T[3]), E = conj(T[4]), F = conj(T[5]);
    const cz U = Zp[0], V = Zp[1], W = Zp[2], X = Zp[3], Y = Zp[4], Z = Zp[5];
    const cz UXY = U + (X + Y), VWZ = V + (W + Z);
    const cz ADE = A + (D + E), BCF = B + (C + F);
    const cz first = A * U + (B * V + (C * W + (D * X + (E * Y + F * Z))));
    const cz second = (UXY - 2.0 * VWZ) * ADE;
    const cz third = (VWZ - 2.0 * UXY) * BCF;
    value = 3.0 * first + (second + third);
  } else {  // get_energy_same<Complex>, Equations.cxx:209-226: cyclic permutations only
    const cz A = conj(T[0]), B = conj(T[3]), C = conj(T[4]);
    const cz U = Zp[0], V = Zp[3], W = Zp[4];
    const cz ABC = A + (B + C), UVW = U + (V + W);
    value = 3.0 * ((A * U + B * V) + C * W) - ABC * UVW;
  }
  return fac * ((2.0 * value) / den);
}

// host restatement of the whole triangular sum over plain [i + j No + k No^2] complex cubes, built
// on point_energy_z (CPU tests only; the device walks tiles instead)
inline double host_energy_z(int No, double epsabc, const double *epsi, const double *Tijk, const double *Zijk,
                            bool same) {
  const size_t N = (size_t)No, NN = N * N;
  auto at = [&](const double *c, size_t x, size_t y, size_t z) {
    const size_t o = 2 * (x + N * y + NN * z);
    return cz{c[o], c[o + 1]};
  };
  double sum = 0.0;
  for (size_t i = 0; i < N; i++)
    for (size_t j = 0; j <= i; j++)
      for (size_t k = 0; k <= j; k++) {
        const cz T[6] = {at(Tijk, i, j, k), at(Tijk, i, k, j), at(Tijk, j, i, k),
                         at(Tijk, j, k, i), at(Tijk, k, i, j), at(Tijk, k, j, i)};
        const cz Z[6] = {at(Zijk, i, j, k), at(Zijk, i, k, j), at(Zijk, j, i, k),
                         at(Zijk, j, k, i), at(Zijk, k, i, j), at(Zijk, k, j, i)};
        const cz eijk = (cz{epsi[2 * i], epsi[2 * i + 1]} + cz{epsi[2 * j], epsi[2 * j + 1]}) +
                        cz{epsi[2 * k], epsi[2 * k + 1]};
        const cz den = cz{epsabc, 0.0} - eijk;
        const double fac = ((j == k) ? 0.5 : 1.0) * ((i == j) ? 0.5 : 1.0);
        sum += point_energy_z(same, T, Z, den, fac).re;
      }
  return sum;
}

__host__ __device__ inline size_t reduce_z_smem_bytes(int No, bool ct) {
  return sizeof(double) * (2 * ((size_t)(ct ? 12 : 6) * RTILE + 18 * 64 + 4 * (size_t)No) + 32);
}

// ReduceParams as in reduction.cuh with: R / RZ = [ntuples][6][cube_stride] (classes re 0-2, im
// 3-5), eps_i / eps_a / Tai / VIJ interleaved complex.
template <bool CT>
__global__ void __launch_bounds__(REDUCE_THREADS, 1)
reduce_z_kernel(const ReduceParams P) {
  extern __shared__ double sm[];
  const int No = P.No, Nv = P.Nv;
  double *WtR = sm, *WtI = WtR + 6 * RTILE;            // Tijk tiles, re / im
  double *ZtR = CT ? WtI + 6 * RTILE : WtR, *ZtI = CT ? ZtR + 6 * RTILE : WtI;  // alias when !CT
  double *VbR = sm + (CT ? 24 : 12) * RTILE, *VbI = VbR + 18 * 64;              // [3][6][64]
  double *sEpsR = VbI + 18 * 64, *sEpsI = sEpsR + No;
  double *sTaR = sEpsI + No, *sTaI = sTaR + No, *sTbR = sTaI + No, *sTbI = sTbR + No, *sTcR = sTbI + No,
         *sTcI = sTcR + No;
  double *sRed = sTcI + No;  // [32]

  const int tup = blockIdx.x;
  const TupleRec rec = P.recs[tup];
  const int tid = threadIdx.x;
  const int split = blockIdx.y;
  if (rec.fake) {
    if (tid == 0) P.e_tuple[(size_t)tup * P.nsplit + split] = 0.0;
    return;
  }
  const int a = rec.a, b = rec.b, c = rec.c;
  const size_t NoNo = (size_t)No * No, cube = P.cube_stride;
  const double *Cr = P.R + (size_t)tup * 6 * cube, *Ci = Cr + 3 * cube;   // 3 classes each
  const double *Zr = P.RZ + (size_t)tup * 6 * cube, *Zi = Zr + 3 * cube;
  const double *Vmat[3];  // VBCij, VACij, VABij (interleaved complex)
#pragma unroll
  for (int q = 0; q < 3; q++)
    Vmat[q] = rec.vij[q] >= P.ownedV ? P.VIJc + 2 * (size_t)(rec.vij[q] - P.ownedV) * NoNo
                                     : P.VIJ + 2 * (size_t)rec.vij[q] * NoNo;
  for (int i = tid; i < No; i += REDUCE_THREADS) {
    sEpsR[i] = P.eps_i[2 * i];
    sEpsI[i] = P.eps_i[2 * i + 1];
    sTaR[i] = P.Tai[2 * (a + (size_t)i * Nv)];
    sTaI[i] = P.Tai[2 * (a + (size_t)i * Nv) + 1];
    sTbR[i] = P.Tai[2 * (b + (size_t)i * Nv)];
    sTbI[i] = P.Tai[2 * (b + (size_t)i * Nv) + 1];
    sTcR[i] = P.Tai[2 * (c + (size_t)i * Nv)];
    sTcI[i] = P.Tai[2 * (c + (size_t)i * Nv) + 1];
  }
  // Atrip.cxx:643-644: epsabc = real(eps_a + eps_b + eps_c)
  const double epsabc = P.eps_a[2 * a] + P.eps_a[2 * b] + P.eps_a[2 * c];
  const bool same = (a == b) != (b == c);

  const int nb = (No + RT - 1) / RT;
  double esum = 0.0;
  const int l0 = tid & 7, l1 = (tid >> 3) & 7, l2 = tid >> 6;

  int orbit = -1;
  for (int I = 0; I < nb; I++)
    for (int J = 0; J <= I; J++)
      for (int K = 0; K <= J; K++) {
        if (++orbit % P.nsplit != split) continue;
        const bool eIJ = (I == J), eJK = (J == K);
        const int c1 = eJK ? 0 : 1, c2 = eIJ ? 0 : 2, c3 = (eIJ && eJK) ? 0 : (eIJ ? 1 : 3),
                  c4 = (eIJ && eJK) ? 0 : (eJK ? 2 : 4), c5 = (eIJ && eJK) ? 0 : (eIJ ? 4 : (eJK ? 3 : 5));
        const int cmap[6] = {0, c1, c2, c3, c4, c5};
        __syncthreads();  // previous orbit fully consumed (and the prologue published)
        for_tiles([&](auto pc) {
          constexpr int p = decltype(pc)::value;
          using T = TilePerm<p>;
          if (cmap[p] == p) {  // distinct tile: build it once
            const size_t tb =
                (((size_t)pick3(T::Z, I, J, K) * nb + pick3(T::Y, I, J, K)) * nb + pick3(T::X, I, J, K)) * 512 + tid;
            double vr[4], vi[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const size_t o = tb + 128 * q;
              vr[q] = (Cr[o] + Cr[o + cube]) + Cr[o + 2 * cube];
              vi[q] = (Ci[o] + Ci[o + cube]) + Ci[o + 2 * cube];
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const int pos = p * RTILE + tile_pos(l0, l1, l2 + 2 * q);
              WtR[pos] = vr[q];
              WtI[pos] = vi[q];
            }
            if (CT) {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const size_t o = tb + 128 * q;
                vr[q] = (Zr[o] + Zr[o + cube]) + Zr[o + 2 * cube];
                vi[q] = (Zi[o] + Zi[o + cube]) + Zi[o + 2 * cube];
              }
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const int pos = p * RTILE + tile_pos(l0, l1, l2 + 2 * q);
                ZtR[pos] = vr[q];
                ZtI[pos] = vi[q];
              }
            }
          }
        });
        // Vabij blocks, as in reduction.cuh: Vb[mat][pair(X,Y)][xl + 8 yl] = Vmat[x + y No]
#pragma unroll
        for (int q = 0; q < 9; q++) {
          const int e = tid + REDUCE_THREADS * q;
          const int mat = e / 384, r = e - mat * 384, pr = r >> 6, xl = r & 7, yl = (r >> 3) & 7;
          const int X = pr >> 1, Y = (pr & 1) ? (X == 2 ? 1 : 2) : (X == 0 ? 1 : 0);
          const int x = pick3(X, I, J, K) * RT + xl, y = pick3(Y, I, J, K) * RT + yl;
          const double *vm = mat == 0 ? Vmat[0] : (mat == 1 ? Vmat[1] : Vmat[2]);
          const bool ok = x < No && y < No;
          VbR[e] = ok ? vm[2 * (x + (size_t)y * No)] : 0.0;
          VbI[e] = ok ? vm[2 * (x + (size_t)y * No) + 1] : 0.0;
        }
        __syncthreads();
        const int il = l0, jl = l1;
        const int i = I * RT + il, j = J * RT + jl;
        if (i < No && j <= i) {
          const int pij = 0 * 64 + il + 8 * jl, pji = 2 * 64 + jl + 8 * il;
          const cz tai{sTaR[i], sTaI[i]}, taj{sTaR[j], sTaI[j]}, tbi{sTbR[i], sTbI[i]}, tbj{sTbR[j], sTbI[j]},
              tci{sTcR[i], sTcI[i]}, tcj{sTcR[j], sTcI[j]};
          const cz eij = cz{sEpsR[i], sEpsI[i]} + cz{sEpsR[j], sEpsI[j]};
          const double facij = (i == j) ? 0.5 : 1.0;
          auto Vbc = [&](int o) { return cz{VbR[o], VbI[o]}; };
          auto Vac = [&](int o) { return cz{VbR[384 + o], VbI[384 + o]}; };
          auto Vab = [&](int o) { return cz{VbR[768 + o], VbI[768 + o]}; };
#pragma unroll 1
          for (int q = 0; q < 4; q++) {
            const int kl = l2 + 2 * q, k = K * RT + kl;
            if (k <= j) {
              const int o[6] = {tile_pos(il, jl, kl), RTILE * c1 + tile_pos(il, kl, jl),
                                RTILE * c2 + tile_pos(jl, il, kl), RTILE * c3 + tile_pos(jl, kl, il),
                                RTILE * c4 + tile_pos(kl, il, jl), RTILE * c5 + tile_pos(kl, jl, il)};
              cz T[6], Z[6];
#pragma unroll
              for (int p = 0; p < 6; p++) {
                T[p] = cz{WtR[o[p]], WtI[o[p]]};
                Z[p] = cz{ZtR[o[p]], ZtI[o[p]]};
              }
              const int pik = 1 * 64 + il + 8 * kl, pjk = 3 * 64 + jl + 8 * kl;
              const int pki = 4 * 64 + kl + 8 * il, pkj = 5 * 64 + kl + 8 * jl;
              const cz tak{sTaR[k], sTaI[k]}, tbk{sTbR[k], sTbI[k]}, tck{sTcR[k], sTcI[k]};
              // Z[x,y,z] = T[x,y,z] + Tai[a,x] Vbc[y,z] + Tai[b,y] Vac[x,z] + Tai[c,z] Vab[x,y]
              Z[0] = ((Z[0] + tai * Vbc(pjk)) + tbj * Vac(pik)) + tck * Vab(pij);  // Z[i,j,k]
              Z[1] = ((Z[1] + tai * Vbc(pkj)) + tbk * Vac(pij)) + tcj * Vab(pik);  // Z[i,k,j]
              Z[2] = ((Z[2] + taj * Vbc(pik)) + tbi * Vac(pjk)) + tck * Vab(pji);  // Z[j,i,k]
              Z[3] = ((Z[3] + taj * Vbc(pki)) + tbk * Vac(pji)) + tci * Vab(pjk);  // Z[j,k,i]
              Z[4] = ((Z[4] + tak * Vbc(pij)) + tbi * Vac(pkj)) + tcj * Vab(pki);  // Z[k,i,j]
              Z[5] = ((Z[5] + tak * Vbc(pji)) + tbj * Vac(pki)) + tci * Vab(pkj);  // Z[k,j,i]
              const cz den = cz{epsabc, 0.0} - (eij + cz{sEpsR[k], sEpsI[k]});
              const double fac = ((j == k) ? 0.5 : 1.0) * facij;
              esum += point_energy_z(same, T, Z, den, fac).re;
            }
          }
        }
      }

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

// debug / parity only: materialise the reference's complex Tijk and Zijk (interleaved, plain
// [i + j No + k No^2] order) for one tuple of the batch
__global__ void cubes_z_kernel(const ReduceParams P, int tup, double *Tijk, double *Zijk) {
  const int No = P.No, Nv = P.Nv;
  const size_t NoNo = (size_t)No * No, n = NoNo * No, cube = P.cube_stride;
  const TupleRec rec = P.recs[tup];
  const double *Cr = P.R + (size_t)tup * 6 * cube, *Ci = Cr + 3 * cube;
  const double *Vm[3];
  for (int q = 0; q < 3; q++)
    Vm[q] = rec.vij[q] >= P.ownedV ? P.VIJc + 2 * (size_t)(rec.vij[q] - P.ownedV) * NoNo
                                   : P.VIJ + 2 * (size_t)rec.vij[q] * NoNo;
  auto ld = [](const double *p, size_t e) { return cz{p[2 * e], p[2 * e + 1]}; };
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % No), j = (int)((e / No) % No), k = (int)(e / NoNo);
    const size_t o = cube_offset(No, i, j, k);
    const cz w{(Cr[o] + Cr[o + cube]) + Cr[o + 2 * cube], (Ci[o] + Ci[o + cube]) + Ci[o + 2 * cube]};
    if (Tijk) {
      Tijk[2 * e] = w.re;
      Tijk[2 * e + 1] = w.im;
    }
    if (Zijk) {
      const cz z = ((w + ld(P.Tai, rec.a + (size_t)i * Nv) * ld(Vm[0], j + (size_t)k * No)) +
                    ld(P.Tai, rec.b + (size_t)j * Nv) * ld(Vm[1], i + (size_t)k * No)) +
                   ld(P.Tai, rec.c + (size_t)k * Nv) * ld(Vm[2], i + (size_t)j * No);
      Zijk[2 * e] = z.re;
      Zijk[2 * e + 1] = z.im;
    }
  }
}

}  // namespace ab
