// HBM stores of the engine and the kernels that build them.
//
// The reference keeps five "slice unions" per rank (Unions.hpp:77-278) in the layouts CTF::slice
// produces.  Here the slices are stored in the layout the contraction kernel's TMA tensor maps
// read directly (DESIGN.md "Data layout"):
//   AX [xslot][m = p + q No][Kp]   kappa <  Nv      : T_x[E,p,q]  = Tabij[x,E,p,q]     (TAPHH)
//                                  Nv <= kappa < Nv+No: -H_x[q,p,L] = -Vijka[q,p,L,x]   (HHHA)
//                                  rest               : 0
//   BY [bslot][r][Kp]              kappa <  Nv      : V_yz[E,r]   = Vabci[y,z,E,r]     (ABPH)
//                                  hole part, slot (y,z), y<z or "normal" diagonal:
//                                                     Tabij[y,z,L,r]                    (TABHH)
//                                  hole part, slot (y,z), y>z or "transposed" diagonal:
//                                                     Tabij[z,y,r,L]
//   VIJ[vslot][i + j No]           Vabij[y,z,i,j]                                       (ABHH)
// Slots are looked up through int tables (xtab, btab, vtab) so that a rank that stores only its
// own slices plus a fetch cache uses the same kernels as a rank that stores everything.
#pragma once
#include "common.cuh"

namespace ab {

struct StoreDims {
  int No, Nv, Kp;
  int cplx = 0;  // 1: F = std::complex<double> (planes along kappa, see "complex field" below)
};

// ------------------------------------------------------------------ synthetic fill (device)
// one thread per store element; the source tensor's column-major linear index is rebuilt and
// hashed, so the stores hold exactly what ingesting oracle_fill()ed host tensors would give.
__global__ void fill_AX_kernel(double *AX, StoreDims d, const int *xlist, int nx, uint64_t keyT, uint64_t keyH,
                               double scale) {
  const size_t per = (size_t)d.No * d.No * d.Kp, total = per * nx;
  const size_t Nv = d.Nv, No = d.No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t s = e / per, r = e - s * per;
    const size_t m = r / d.Kp, kap = r - m * d.Kp;
    const size_t p = m % No, q = m / No, x = xlist[s];
    double v = 0.0;
    if (kap < Nv) v = synth_val(keyT, x + kap * Nv + p * Nv * Nv + q * Nv * Nv * No, scale);
    else if (kap < Nv + No) {
      const size_t L = kap - Nv;
      v = -synth_val(keyH, q + p * No + L * No * No + x * No * No * No, scale);
    }
    AX[e] = v;
  }
}

// slot s holds ordered pair (ylist[s], zlist[s]); tflag[s] != 0 marks the transposed diagonal
__global__ void fill_BY_kernel(double *BY, StoreDims d, const int *ylist, const int *zlist, const int *tflag,
                               size_t nb, uint64_t keyV, uint64_t keyT, double scale) {
  const size_t per = (size_t)d.No * d.Kp, total = per * nb;
  const size_t Nv = d.Nv, No = d.No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t s = e / per, rr = e - s * per;
    const size_t r = rr / d.Kp, kap = rr - r * d.Kp;
    const size_t y = ylist[s], z = zlist[s];
    double v = 0.0;
    if (kap < Nv) v = synth_val(keyV, y + z * Nv + kap * Nv * Nv + r * Nv * Nv * Nv, scale);
    else if (kap < Nv + No) {
      const size_t L = kap - Nv;
      const bool transposed = (y > z) || (y == z && tflag[s]);
      v = transposed ? synth_val(keyT, z + y * Nv + r * Nv * Nv + L * Nv * Nv * No, scale)
                     : synth_val(keyT, y + z * Nv + L * Nv * Nv + r * Nv * Nv * No, scale);
    }
    BY[e] = v;
  }
}

__global__ void fill_VIJ_kernel(double *VIJ, StoreDims d, const int *ylist, const int *zlist, size_t nv,
                                uint64_t key, double scale) {
  const size_t per = (size_t)d.No * d.No, total = per * nv;
  const size_t Nv = d.Nv, No = d.No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t s = e / per, r = e - s * per;
    const size_t i = r % No, j = r / No;
    VIJ[e] = synth_val(key, (size_t)ylist[s] + (size_t)zlist[s] * Nv + i * Nv * Nv + j * Nv * Nv * No, scale);
  }
}

__global__ void fill_small_kernel(double *eps_i, double *eps_a, double *Tai, StoreDims d, uint64_t kI, uint64_t kA,
                                  uint64_t kT, double scale) {
  const size_t n = (size_t)d.No + d.Nv + (size_t)d.No * d.Nv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    if (e < (size_t)d.No) eps_i[e] = __dadd_rn(-2.0, __dmul_rn(1.5, synth_u(kI, e)));
    else if (e < (size_t)d.No + d.Nv) eps_a[e - d.No] = __dadd_rn(0.5, __dmul_rn(3.5, synth_u(kA, e - d.No)));
    else Tai[e - d.No - d.Nv] = synth_val(kT, e - d.No - d.Nv, scale);
  }
}

// ------------------------------------------------------------------ ingest from host tensors
// Each kernel consumes one contiguous chunk of a CTF-layout tensor that was copied to the device
// and scatters it into the stores.  *_tab < 0 means "this rank does not store that slice".

// chunk = Tabij[:, :, p, q] (Nv x Nv, alpha fastest).  Writes
//   AX[xtab[alpha]][p + q No][beta]                       (TAPHH, Unions.hpp:77-113)
//   BY[btab[alpha + beta Nv]][q][Nv + p]         alpha <= beta   (TABHH normal,  Unions.hpp:239-278)
//   BY[btab[beta + alpha Nv] or diagT][p][Nv + q] alpha <= beta   (TABHH transposed)
__global__ void ingest_Tabij_kernel(const double *chunk, StoreDims d, int p, int q, double *AX, const int *xtab,
                                    double *BY, const int *btab) {
  __shared__ double tile[32][33];
  const int Nv = d.Nv, No = d.No;
  const size_t Kp = d.Kp;
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int al = a0 + threadIdx.x, be = b0 + r;
    double v = 0.0;
    if (al < Nv && be < Nv) {
      v = chunk[al + (size_t)be * Nv];
      if (al <= be) {
        const int s1 = btab[al + be * Nv];
        if (s1 >= 0) BY[((size_t)s1 * No + q) * Kp + Nv + p] = v;
        const int s2 = btab[al == be ? Nv * Nv + al : be + al * Nv];
        if (s2 >= 0) BY[((size_t)s2 * No + p) * Kp + Nv + q] = v;
      }
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int al = a0 + r, be = b0 + threadIdx.x;
    if (al < Nv && be < Nv) {
      const int s = xtab[al];
      if (s >= 0) AX[((size_t)s * No * No + p + (size_t)q * No) * Kp + be] = tile[threadIdx.x][r];
    }
  }
}

// chunk = Vijka[:, :, :, x0 .. x0+nx) (No^3 per x).  AX[xtab[x]][p + q No][Nv + L] = -Vijka[q,p,L,x]
// (HHHA, Unions.hpp:115-152; sign and (p,q) swap folded in here, see contraction.cuh)
__global__ void ingest_Vijka_kernel(const double *chunk, StoreDims d, int x0, int nx, double *AX, const int *xtab) {
  const size_t No = d.No, cube = No * No * No, total = cube * nx;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t xs = e / cube, r = e - xs * cube;
    // enumerate destination-friendly: L fastest
    const size_t L = r % No, m = r / No, p = m % No, q = m / No;
    const int s = xtab[x0 + xs];
    if (s >= 0) AX[((size_t)s * No * No + m) * d.Kp + d.Nv + L] = -chunk[xs * cube + q + p * No + L * No * No];
  }
}

// chunk = Vabij[:, :, i, j] (Nv x Nv).  VIJ[vtab[y + z Nv]][i + j No]   (ABHH, Unions.hpp:199-237)
__global__ void ingest_Vabij_kernel(const double *chunk, StoreDims d, int i, int j, double *VIJ, const int *vtab) {
  const size_t n = (size_t)d.Nv * d.Nv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int s = vtab[e];
    if (s >= 0) VIJ[(size_t)s * d.No * d.No + i + (size_t)j * d.No] = chunk[e];
  }
}

// chunk = Vabci[:, :, E0 .. E0+ne, r] (ne x Nv^2, pair index fastest).  BY[btab[pair]][r][E0 + e]
// and the transposed-diagonal twin of (y,y)   (ABPH, Unions.hpp:154-197)
__global__ void ingest_Vabci_kernel(const double *chunk, StoreDims d, int E0, int ne, int r, double *BY,
                                    const int *btab) {
  __shared__ double tile[16][33];
  const size_t NvNv = (size_t)d.Nv * d.Nv;
  const size_t pr0 = (size_t)blockIdx.x * 32;
  // load: threadIdx.x runs along the pair index (coalesced), threadIdx.y along E
  for (int e = threadIdx.y; e < 16; e += blockDim.y) {
    const size_t pr = pr0 + threadIdx.x;
    tile[e][threadIdx.x] = (e < ne && pr < NvNv) ? chunk[(size_t)e * NvNv + pr] : 0.0;
  }
  __syncthreads();
  // store: 16 consecutive E per pair
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int w = tid; w < 32 * 16; w += blockDim.x * blockDim.y) {
    const int e = w & 15, pl = w >> 4;
    const size_t pr = pr0 + pl;
    if (pr < NvNv && e < ne) {
      const int s = btab[pr];
      const double v = tile[e][pl];
      if (s >= 0) BY[((size_t)s * d.No + r) * d.Kp + E0 + e] = v;
      const int y = (int)(pr % d.Nv), z = (int)(pr / d.Nv);
      if (y == z) {
        const int s2 = btab[NvNv + y];
        if (s2 >= 0) BY[((size_t)s2 * d.No + r) * d.Kp + E0 + e] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ complex field (F = Complex)
// The complex contraction C = A B (A = [T_x ; -conj(H_x)], B = [V_yz ; T_yz], reference
// Equations.cxx:620-680 with MAYBE_CONJ on the hole integrals, :623-648) runs on the SAME real
// DMMA kernel with the contraction index doubled, Kh = No + Nv:
//     Re C = [ Re A | -Im A ] [ Re B ; Im B ]          Im C = [ Im A | Re A ] [ Re B ; Im B ]
//   AX [xslot][variant][m][Kp]   variant 0: kappa < Kh: Re A[kappa];  Kh <= kappa < 2Kh: -Im A[kappa - Kh]
//                                variant 1: kappa < Kh: Im A[kappa];  Kh <= kappa < 2Kh:  Re A[kappa - Kh]
//   BY [bslot][r][Kp]            kappa < Kh: Re B[kappa];             Kh <= kappa < 2Kh:  Im B[kappa - Kh]
//   VIJ[vslot][i + j No]         interleaved (re, im)
// with A[kk] = T_x[kk,p,q] (kk < Nv) or -conj(H_x[q,p,kk-Nv]), B[kk] as in the real case, and
// Kp = 2 Kh rounded up to 16.  Both variants of a slice are contiguous, so the slice exchange moves
// them as one slice.  Synthetic complex tensors: element e = (synth(2e), synth(2e+1)).

// what a store element holds: part (0 re, 1 im) of element `lin` of source tensor `tensor`
// (0 = padding, 1 = Tabij, 2 = Vijka/Jijka, 3 = Vabci/Jabci), times `sign`
struct SourceRef {
  int tensor;
  int part;
  double sign;
  unsigned long long lin;
};

__host__ __device__ inline SourceRef ax_source_z(int No_, int Nv_, int variant, size_t x, size_t m, size_t kap) {
  const size_t No = No_, Nv = Nv_, Kh = No + Nv;
  SourceRef r{0, 0, 0.0, 0ull};
  if (kap >= 2 * Kh) return r;
  const int plane = kap >= Kh;
  const size_t kk = plane ? kap - Kh : kap;
  // (Re A, Im A) of this kk; variant 0 wants (Re, -Im), variant 1 wants (Im, Re)
  const int want_im = (variant == 0) ? plane : !plane;
  const double vsign = (variant == 0 && plane) ? -1.0 : 1.0;
  const size_t p = m % No, q = m / No;
  if (kk < Nv) {  // A = T_x[E,p,q]
    r.tensor = 1;
    r.lin = x + kk * Nv + p * Nv * Nv + q * Nv * Nv * No;
    r.part = want_im;
    r.sign = vsign;
  } else {  // A = -conj(H_x[q,p,L]): Re A = -Re H, Im A = +Im H
    const size_t L = kk - Nv;
    r.tensor = 2;
    r.lin = q + p * No + L * No * No + x * No * No * No;
    r.part = want_im;
    r.sign = want_im ? vsign : -vsign;
  }
  return r;
}

__host__ __device__ inline SourceRef by_source_z(int No_, int Nv_, size_t y, size_t z, int tflag, size_t rr,
                                                 size_t kap) {
  const size_t No = No_, Nv = Nv_, Kh = No + Nv;
  SourceRef r{0, 0, 0.0, 0ull};
  if (kap >= 2 * Kh) return r;
  r.part = kap >= Kh;
  r.sign = 1.0;
  const size_t kk = r.part ? kap - Kh : kap;
  if (kk < Nv) {
    r.tensor = 3;
    r.lin = y + z * Nv + kk * Nv * Nv + rr * Nv * Nv * Nv;
  } else {
    const size_t L = kk - Nv;
    const bool transposed = (y > z) || (y == z && tflag);
    r.tensor = 1;
    r.lin = transposed ? z + y * Nv + rr * Nv * Nv + L * Nv * Nv * No : y + z * Nv + L * Nv * Nv + rr * Nv * Nv * No;
  }
  return r;
}

__global__ void fill_AX_z_kernel(double *AX, StoreDims d, const int *xlist, int nx, uint64_t keyT, uint64_t keyH,
                                 double scale) {
  const size_t per = (size_t)d.No * d.No * d.Kp, total = per * 2 * nx;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t sv = e / per, r = e - sv * per;
    const size_t s = sv >> 1, m = r / d.Kp, kap = r - m * d.Kp;
    const SourceRef sr = ax_source_z(d.No, d.Nv, (int)(sv & 1), (size_t)xlist[s], m, kap);
    double v = 0.0;
    if (sr.tensor) v = __dmul_rn(sr.sign, synth_val(sr.tensor == 1 ? keyT : keyH, 2 * sr.lin + sr.part, scale));
    AX[e] = v;
  }
}

__global__ void fill_BY_z_kernel(double *BY, StoreDims d, const int *ylist, const int *zlist, const int *tflag,
                                 size_t nb, uint64_t keyV, uint64_t keyT, double scale) {
  const size_t per = (size_t)d.No * d.Kp, total = per * nb;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t s = e / per, rr = e - s * per;
    const size_t r = rr / d.Kp, kap = rr - r * d.Kp;
    const SourceRef sr = by_source_z(d.No, d.Nv, (size_t)ylist[s], (size_t)zlist[s], tflag[s], r, kap);
    double v = 0.0;
    if (sr.tensor) v = synth_val(sr.tensor == 3 ? keyV : keyT, 2 * sr.lin + sr.part, scale);
    BY[e] = v;
  }
}

// VIJ (interleaved complex) and Tai: plain doubled index; eps: (real-case value, 0)
__global__ void fill_VIJ_z_kernel(double *VIJ, StoreDims d, const int *ylist, const int *zlist, size_t nv,
                                  uint64_t key, double scale) {
  const size_t per = 2 * (size_t)d.No * d.No, total = per * nv;
  const size_t Nv = d.Nv, No = d.No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t s = e / per, r2 = e - s * per, r = r2 >> 1;
    const size_t i = r % No, j = r / No;
    const size_t lin = (size_t)ylist[s] + (size_t)zlist[s] * Nv + i * Nv * Nv + j * Nv * Nv * No;
    VIJ[e] = synth_val(key, 2 * lin + (r2 & 1), scale);
  }
}

__global__ void fill_small_z_kernel(double *eps_i, double *eps_a, double *Tai, StoreDims d, uint64_t kI, uint64_t kA,
                                    uint64_t kT, double scale) {
  const size_t n = (size_t)d.No + d.Nv + (size_t)d.No * d.Nv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    if (e < (size_t)d.No) {
      eps_i[2 * e] = __dadd_rn(-2.0, __dmul_rn(1.5, synth_u(kI, e)));
      eps_i[2 * e + 1] = 0.0;
    } else if (e < (size_t)d.No + d.Nv) {
      const size_t a = e - d.No;
      eps_a[2 * a] = __dadd_rn(0.5, __dmul_rn(3.5, synth_u(kA, a)));
      eps_a[2 * a + 1] = 0.0;
    } else {
      const size_t t = e - d.No - d.Nv;
      Tai[2 * t] = synth_val(kT, 2 * t, scale);
      Tai[2 * t + 1] = synth_val(kT, 2 * t + 1, scale);
    }
  }
}

// ---- ingest of interleaved-complex host chunks (one thread per source element; the same chunking
//      as the real kernels above, chunk element e at doubles [2e, 2e+1])

// chunk = Tabij[:, :, p, q]
__global__ void ingest_Tabij_z_kernel(const double *chunk, StoreDims d, int p, int q, double *AX, const int *xtab,
                                      double *BY, const int *btab) {
  const size_t Nv = d.Nv, No = d.No, Kp = d.Kp, Kh = No + Nv, n = Nv * Nv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const size_t al = e % Nv, be = e / Nv;
    const double tr = chunk[2 * e], ti = chunk[2 * e + 1];
    const int s = xtab[al];
    if (s >= 0) {  // A[kk = be] = T_al[be,p,q]
      double *row0 = AX + (((size_t)s * 2 + 0) * No * No + p + (size_t)q * No) * Kp;
      double *row1 = AX + (((size_t)s * 2 + 1) * No * No + p + (size_t)q * No) * Kp;
      row0[be] = tr;
      row0[Kh + be] = -ti;
      row1[be] = ti;
      row1[Kh + be] = tr;
    }
    if (al <= be) {  // hole part of B, as in ingest_Tabij_kernel
      const int s1 = btab[al + be * Nv];
      if (s1 >= 0) {
        double *row = BY + ((size_t)s1 * No + q) * Kp;
        row[Nv + p] = tr;
        row[Kh + Nv + p] = ti;
      }
      const int s2 = btab[al == be ? Nv * Nv + al : be + al * Nv];
      if (s2 >= 0) {
        double *row = BY + ((size_t)s2 * No + p) * Kp;
        row[Nv + q] = tr;
        row[Kh + Nv + q] = ti;
      }
    }
  }
}

// chunk = Vijka[:, :, :, x0 .. x0+nx): A[kk = Nv + L] = -conj(Vijka[q,p,L,x])
__global__ void ingest_Vijka_z_kernel(const double *chunk, StoreDims d, int x0, int nx, double *AX, const int *xtab) {
  const size_t No = d.No, Nv = d.Nv, Kh = No + Nv, cube = No * No * No, total = cube * nx;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t xs = e / cube, r = e - xs * cube;
    const size_t L = r % No, m = r / No, p = m % No, q = m / No;
    const int s = xtab[x0 + xs];
    if (s < 0) continue;
    const size_t src = xs * cube + q + p * No + L * No * No;
    const double hr = chunk[2 * src], hi = chunk[2 * src + 1];  // Re A = -hr, Im A = +hi
    double *row0 = AX + (((size_t)s * 2 + 0) * No * No + m) * d.Kp;
    double *row1 = AX + (((size_t)s * 2 + 1) * No * No + m) * d.Kp;
    row0[Nv + L] = -hr;
    row0[Kh + Nv + L] = -hi;
    row1[Nv + L] = hi;
    row1[Kh + Nv + L] = -hr;
  }
}

// chunk = Vabij[:, :, i, j]
__global__ void ingest_Vabij_z_kernel(const double *chunk, StoreDims d, int i, int j, double *VIJ, const int *vtab) {
  const size_t n = (size_t)d.Nv * d.Nv;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int s = vtab[e];
    if (s < 0) continue;
    double *dst = VIJ + 2 * ((size_t)s * d.No * d.No + i + (size_t)j * d.No);
    dst[0] = chunk[2 * e];
    dst[1] = chunk[2 * e + 1];
  }
}

// chunk = Vabci[:, :, E0 .. E0+ne, r]
__global__ void ingest_Vabci_z_kernel(const double *chunk, StoreDims d, int E0, int ne, int r, double *BY,
                                      const int *btab) {
  const size_t Nv = d.Nv, No = d.No, Kh = No + Nv, NvNv = Nv * Nv, total = NvNv * ne;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t pr = e % NvNv, el = e / NvNv;
    const double vr = chunk[2 * e], vi = chunk[2 * e + 1];
    const int s = btab[pr];
    if (s >= 0) {
      double *row = BY + ((size_t)s * No + r) * d.Kp;
      row[E0 + el] = vr;
      row[Kh + E0 + el] = vi;
    }
    const size_t y = pr % Nv, z = pr / Nv;
    if (y == z) {
      const int s2 = btab[NvNv + y];
      if (s2 >= 0) {
        double *row = BY + ((size_t)s2 * No + r) * d.Kp;
        row[E0 + el] = vr;
        row[Kh + E0 + el] = vi;
      }
    }
  }
}

// read back a slice in the reference layout, interleaved complex; same kinds as read_slice_kernel.
// AXx points at the slice's variant 0 ([Re A | -Im A] rows); VIJxy is interleaved.
__global__ void read_slice_z_kernel(int kind, StoreDims d, const double *AXx, const double *BYxy,
                                    const double *VIJxy, int y, double *out) {
  const size_t No = d.No, Nv = d.Nv, Kp = d.Kp, Kh = No + Nv;
  size_t n = 0;
  if (kind == 100) n = Nv * No * No;
  else if (kind == 101) n = No * No * No;
  else if (kind == 200) n = Nv * No;
  else n = No * No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    double re, im;
    if (kind == 100) {  // TX[E + p Nv + q Nv No] = A[m][E]
      const size_t E = e % Nv, m = e / Nv;
      re = AXx[m * Kp + E];
      im = -AXx[m * Kp + Kh + E];
    } else if (kind == 101) {  // HX[p + q No + L No^2]: A[(q + p No)][Nv + L] = -conj(H)
      const size_t p = e % No, q = (e / No) % No, L = e / (No * No);
      re = -AXx[(q + p * No) * Kp + Nv + L];
      im = -AXx[(q + p * No) * Kp + Kh + Nv + L];  // plane 1 holds -Im A = -Im H
    } else if (kind == 200) {
      const size_t E = e % Nv, r = e / Nv;
      re = BYxy[r * Kp + E];
      im = BYxy[r * Kp + Kh + E];
    } else if (kind == 201) {
      re = AXx[e * Kp + y];
      im = -AXx[e * Kp + Kh + y];
    } else {
      re = VIJxy[2 * e];
      im = VIJxy[2 * e + 1];
    }
    out[2 * e] = re;
    out[2 * e + 1] = im;
  }
}


// ------------------------------------------------------------------ per-slice ingest / read-back
// The reference's SliceUnion<F>::init slices, per rank, only the sources that rank owns
// (SliceUnion.cxx:305-332, Unions.hpp:21-75) and keeps them as contiguous buffers in the slice
// layout CTF::slice yields.  upload_slices_kernel takes a chunk of such buffers (n slices of one
// kind, back to back) and writes them into the stores; slices this rank does not hold are skipped.
//   kind 100 TA(x)      [E + Nv (p + q No)]      = Tabij[x,E,p,q]   -> AX particle part
//        101 VIJKA(x)   [i + j No + k No^2]      = Vijka[i,j,k,x]   -> AX hole part (sign, (p,q) swap)
//        200 VABCI(x,y) [E + Nv r]               = Vabci[x,y,E,r]   -> BY particle part (+ diagonal twin)
//        201 TABIJ(x,y) [p + q No], x <= y       = Tabij[x,y,p,q]   -> BY hole part of (x,y) and of (y,x)'
//        202 VABIJ(x,y) [i + j No], x <= y       = Vabij[x,y,i,j]   -> VIJ
// Z = complex field: source elements are interleaved (re, im), stores as in "complex field" above.
struct SliceTables {
  const int *xtab, *btab, *vtab;  // global id -> owned slot or -1
};

template <bool Z>
__device__ __forceinline__ void put_A(double *AX, const StoreDims &d, size_t slot, size_t m, size_t kk, double re, double im) {
  const size_t No = d.No, Kp = d.Kp;
  if (!Z) {
    AX[(slot * No * No + m) * Kp + kk] = re;
  } else {
    const size_t Kh = (size_t)d.No + d.Nv;
    double *row0 = AX + ((slot * 2 + 0) * No * No + m) * Kp, *row1 = AX + ((slot * 2 + 1) * No * No + m) * Kp;
    row0[kk] = re;
    row0[Kh + kk] = -im;
    row1[kk] = im;
    row1[Kh + kk] = re;
  }
}
template <bool Z>
__device__ __forceinline__ void put_B(double *BY, const StoreDims &d, size_t slot, size_t r, size_t kk, double re, double im) {
  double *row = BY + (slot * d.No + r) * d.Kp;
  row[kk] = re;
  if (Z) row[(size_t)d.No + d.Nv + kk] = im;
}

template <bool Z>
__global__ void upload_slices_kernel(int kind, const double *src, const long long *xy, int n, StoreDims d, SliceTables tb,
                                     double *AX, double *BY, double *VIJ) {
  const size_t No = d.No, Nv = d.Nv;
  const size_t per = kind == 100 ? Nv * No * No : (kind == 101 ? No * No * No : (kind == 200 ? Nv * No : No * No));
  const size_t total = per * (size_t)n;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t s = g / per, e = g - s * per;
    const long long x = xy[2 * s], y = xy[2 * s + 1];
    if (kind == 101) {
      // destination-friendly enumeration (L fastest): AX[m = p + q No][Nv + L] = -conj(Vijka[q,p,L,x])
      const int slot = tb.xtab[x];
      if (slot < 0) continue;
      const size_t L = e % No, m = e / No, p = m % No, q = m / No;
      const size_t si = s * per + q + p * No + L * No * No;
      const double re = Z ? src[2 * si] : src[si], im = Z ? src[2 * si + 1] : 0.0;
      put_A<Z>(AX, d, (size_t)slot, m, Nv + L, -re, im);
      continue;
    }
    const double re = Z ? src[2 * g] : src[g], im = Z ? src[2 * g + 1] : 0.0;
    if (kind == 100) {
      const int slot = tb.xtab[x];
      if (slot >= 0) put_A<Z>(AX, d, (size_t)slot, e / Nv, e % Nv, re, im);
    } else if (kind == 200) {
      const size_t E = e % Nv, r = e / Nv;
      const int s1 = tb.btab[x + y * (long long)Nv];
      if (s1 >= 0) put_B<Z>(BY, d, (size_t)s1, r, E, re, im);
      if (x == y) {
        const int s2 = tb.btab[Nv * Nv + x];
        if (s2 >= 0) put_B<Z>(BY, d, (size_t)s2, r, E, re, im);
      }
    } else if (kind == 201) {
      const size_t p = e % No, q = e / No;
      const int s1 = tb.btab[x + y * (long long)Nv];  // (x,y): hole[r = q][L = p]
      if (s1 >= 0) put_B<Z>(BY, d, (size_t)s1, q, Nv + p, re, im);
      const int s2 = tb.btab[x == y ? (long long)(Nv * Nv) + x : y + x * (long long)Nv];  // (y,x)': hole[r = p][L = q]
      if (s2 >= 0) put_B<Z>(BY, d, (size_t)s2, p, Nv + q, re, im);
    } else {  // 202
      const int slot = tb.vtab[x + y * (long long)Nv];
      if (slot >= 0) {
        if (Z) {
          VIJ[2 * ((size_t)slot * No * No + e)] = re;
          VIJ[2 * ((size_t)slot * No * No + e) + 1] = im;
        } else {
          VIJ[(size_t)slot * No * No + e] = re;
        }
      }
    }
  }
}

// n slices of one kind back to the reference's slice layout (the inverse of the above); a slice
// this rank does not hold comes back as NaN.  TABIJ(x,y) is read from the TA(x) rows when x is owned
// (as read_slice_kernel does), else from the hole part of the (y,x)' slot held for y.
template <bool Z>
__global__ void read_slices_kernel(int kind, double *out, const long long *xy, int n, StoreDims d, SliceTables tb,
                                   const double *AX, const double *BY, const double *VIJ) {
  const size_t No = d.No, Nv = d.Nv, Kp = d.Kp, Kh = No + Nv;
  const size_t per = kind == 100 ? Nv * No * No : (kind == 101 ? No * No * No : (kind == 200 ? Nv * No : No * No));
  const size_t total = per * (size_t)n;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t s = g / per, e = g - s * per;
    const long long x = xy[2 * s], y = xy[2 * s + 1];
    double re = nan, im = nan;
    if (kind == 201 && tb.xtab[x] < 0) {
      // x is not owned here: the slice lives in the hole part of the (y,x)' slot this rank holds for y
      const int s2 = tb.btab[x == y ? (long long)(Nv * Nv) + x : y + x * (long long)Nv];
      if (s2 >= 0) {
        const size_t p = e % No, q = e / No;
        const double *row = BY + ((size_t)s2 * No + p) * Kp;
        re = row[Nv + q];
        if (Z) im = row[Kh + Nv + q];
      }
    } else if (kind == 100 || kind == 101 || kind == 201) {
      const int slot = tb.xtab[x];
      if (slot >= 0) {
        const double *A = AX + (size_t)slot * (Z ? 2 : 1) * No * No * Kp;  // variant 0 rows: [Re A | -Im A]
        size_t m, kk;
        double sg = 1.0;
        if (kind == 100) { m = e / Nv; kk = e % Nv; }
        else if (kind == 201) { m = e; kk = (size_t)y; }
        else { const size_t p = e % No, q = (e / No) % No, L = e / (No * No); m = q + p * No; kk = Nv + L; sg = -1.0; }
        re = sg * A[m * Kp + kk];
        // hole part holds -conj(H): Re A = -Re H, Im A = +Im H; plane 1 of variant 0 holds -Im A
        if (Z) im = -A[m * Kp + Kh + kk];
      }
    } else if (kind == 200) {
      const int slot = tb.btab[x + y * (long long)Nv];
      if (slot >= 0) {
        const size_t E = e % Nv, r = e / Nv;
        const double *row = BY + ((size_t)slot * No + r) * Kp;
        re = row[E];
        if (Z) im = row[Kh + E];
      }
    } else {
      const int slot = tb.vtab[x + y * (long long)Nv];
      if (slot >= 0) {
        re = Z ? VIJ[2 * ((size_t)slot * No * No + e)] : VIJ[(size_t)slot * No * No + e];
        if (Z) im = VIJ[2 * ((size_t)slot * No * No + e) + 1];
      }
    }
    if (Z) { out[2 * g] = re; out[2 * g + 1] = im; }
    else out[g] = re;
  }
}

// ------------------------------------------------------------------ read back in reference layout
// kind: 100 TA(x), 101 VIJKA(x), 200 VABCI(x,y), 201 TABIJ(x,y), 202 VABIJ(x,y)
__global__ void read_slice_kernel(int kind, StoreDims d, const double *AXx, const double *BYxy, const double *VIJxy,
                                  int y, double *out) {
  const size_t No = d.No, Nv = d.Nv, Kp = d.Kp;
  size_t n = 0;
  if (kind == 100) n = Nv * No * No;
  else if (kind == 101) n = No * No * No;
  else if (kind == 200) n = Nv * No;
  else n = No * No;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    if (kind == 100) {  // TX[E + p Nv + q Nv No]
      const size_t E = e % Nv, m = e / Nv;
      out[e] = AXx[m * Kp + E];
    } else if (kind == 101) {  // HX[p + q No + L No^2] = -AX[(q + p No)][Nv + L]
      const size_t p = e % No, q = (e / No) % No, L = e / (No * No);
      out[e] = -AXx[(q + p * No) * Kp + Nv + L];
    } else if (kind == 200) {  // VXY[E + r Nv]
      const size_t E = e % Nv, r = e / Nv;
      out[e] = BYxy[r * Kp + E];
    } else if (kind == 201) {  // TXY[p + q No] = T_x[E = y, p, q]
      out[e] = AXx[e * Kp + y];
    } else {
      out[e] = VIJxy[e];
    }
  }
}

}  // namespace ab
