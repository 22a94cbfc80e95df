// Kernel 1 of the (T) hot path: the doubles contraction.
//
// Replaces doubles_contribution (reference Equations.cxx:455-728): 12 xgemm calls
// (Blas.cxx:46-130) + 12 reorder<perm> accumulations (Equations.cxx:26-83) + 3 copies +
// 3 memsets per tuple.  Here one persistent launch processes a whole batch of tuples, and per
// tuple the twelve terms collapse to THREE FP64 tensor-core GEMMs of shape
// [No^2 x 2Kp] x [2Kp x No] (Kp = Nv + No padded to 16), because with
//     A_x[(p,q), kappa] = [ T_x[E,p,q] ; -H_x[q,p,L] ]            (kappa = E, then Nv + L)
//     B_yz[kappa, r]    = [ V_yz[E,r] ; T_yz[L,r] or T_zy[r,L] ]
// the reference's terms are
//     C_k[i + j No, k] = A_a [(i,j),:] B_bc + A_b^T[(i,j),:] B_ac      (P1 + H5, P5 + H3)
//     C_j[i + k No, j] = A_a [(i,k),:] B_cb + A_c^T[(i,k),:] B_ab      (P2 + H6, P3 + H1)
//     C_i[j + k No, i] = A_b [(j,k),:] B_ca + A_c^T[(j,k),:] B_ba      (P6 + H4, P4 + H2)
//     Tijk[i,j,k]      = C_k[i,j,k] + C_j[i,k,j] + C_i[j,k,i]
// (the epilogue stores each class cube at Tijk's own [i + j No + k No^2], so kernel 2 just adds)
// with A^T[(u,v),:] = A[(v,u),:].  The "transposition" is only a second TMA tensor map over the
// same HBM store with the two row strides exchanged -- the permutations live in the operand
// index maps, there is no reorder pass and no scratch GEMM output that gets re-accumulated.
//
// Structure (per CTA, persistent over work items = (tuple, class, row tile, column tile)):
//   warp NW        : producer; one lane issues cp.async.bulk.tensor (TMA, SWIZZLE_128B) for the
//                    A tile [tu*tv rows x 16] and the B tile [<=NI*8 rows x 16] of each K chunk
//                    into an nstages-deep ring guarded by full/empty mbarriers; it runs ahead
//                    across work items so the pipe never drains between tiles.
//   warps 0..NW-1  : consumers; each owns MI x NI DMMA.8x8x4 accumulator fragments
//                    (rows [w*MI*8, (w+1)*MI*8) of the tile), loads fragments with LDS.64 from
//                    the swizzled rows (conflict free: fragment row g <-> tile row 2(g%4)+g/4),
//                    and stores its C fragments straight to the class cube in HBM/L2.
#pragma once
#include "common.cuh"
#include "schedule.hpp"

namespace ab {

constexpr int MAX_STAGES = 12;

// Storage of a class cube: 8x8x8 tiles, each 4 KB contiguous ([x + 8 y + 64 z] inside the tile),
// tiles ordered [X + nb Y + nb^2 Z], nb = ceil(No/8).  Kernel 2 reads whole tiles (full 128-byte
// lines; a plain [i + j No + k No^2] cube costs it 2x the DRAM traffic, ncu r01b).
__host__ __device__ inline size_t cube_blocked_elems(int No) {
  const size_t nb = (size_t)(No + 7) / 8;
  return nb * nb * nb * 512;
}
__host__ __device__ inline size_t cube_offset(int No, int i, int j, int k) {
  const size_t nb = (size_t)(No + 7) / 8;
  return (((size_t)(k >> 3) * nb + (j >> 3)) * nb + (i >> 3)) * 512 + (i & 7) + 8 * (j & 7) + 64 * (k & 7);
}

// tensor maps of the owned stores and of the fetch caches (same layouts; a rank that stores
// everything passes the owned maps twice)
struct ContractMaps {
  CUtensorMap A, AT, B;     // owned AX (plain / rows transposed), owned BY
  CUtensorMap Ac, ATc, Bc;  // cache AX (plain / transposed), cache BY
};

struct ContractParams {
  int No, Nv, Kp;
  int nk;       // Kp / KC: K chunks per operand pair (a class runs 2*nk chunks)
  int last_steps;  // k-steps (of 4) with data in the last chunk: ceil((No + Nv - (nk-1) KC) / 4)
  int tu, tv;   // row tile = tu values of u (fast) x tv values of v  ->  tu*tv rows (u + v No = row of C)
  int utiles;   // ceil(No / tu)
  int mtiles;   // utiles * ceil(No / tv)
  int ntiles;   // ceil(No / (NI*8))
  int arows;    // shared-memory rows reserved for A per stage = NW*MI*8 >= tu*tv
  int brows;    // rows of the B TMA box (min(NI*8, No))
  int nstages;
  int ntuples;
  int ownedA, ownedB;      // slots >= owned address the cache maps (schedule.hpp)
  const TupleRec *recs;    // the batch: tuple + store slots of its slices (built on the host)
  double *R;               // class cubes C_k, C_j, C_i (8x8x8-blocked) of tuple t at R + t tuple_stride
  size_t cube_stride;      // doubles per stored cube = ceil(No/8)^3 * 512
  size_t tuple_stride;     // 3 cube_stride; complex field: 6 cube_stride (Re cubes, then Im cubes --
                           // the Im launch gets R + 3 cube_stride and the variant-1 tensor maps)
};

__host__ __device__ inline size_t contract_stage_bytes(int arows, int NI) {
  return (size_t)(arows + NI * 8) * 128;
}

// CREGS > 0: register re-allocation between the warp roles (setmaxnreg).  The block then carries a
// whole producer warpgroup behind the consumer warps (warps nwarps .. nwarps+3, only the first one
// works): it shrinks to 24 registers per thread and the consumer warpgroups grow to CREGS, beyond
// the 65536 / MAXT cap of a uniform allocation (the register file is split per SM sub-partition: with
// 8 consumer warps + 1 producer warp one sub-partition hosts 3 warps, which caps every thread at 168
// registers and made the 4x8-fragment variant spill).  Needs nwarps % 4 == 0.
template <int MI, int NI, int MAXT, int CREGS = 0>
__global__ void __launch_bounds__(MAXT, 1)
contract_kernel(const __grid_constant__ ContractMaps M, const ContractParams P) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES];

  const int nwarps = (blockDim.x >> 5) - (CREGS ? 4 : 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const size_t stage_bytes = contract_stage_bytes(P.arows, NI);

  // zero the ring once: rows the TMA boxes never write (B rows >= brows, A rows >= No*tv) must
  // not hold NaN patterns from a previous kernel
  {
    const size_t n16 = stage_bytes * P.nstages / 16;
    uint4 *z = reinterpret_cast<uint4 *>(base);
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.nstages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], nwarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&M.A);
    tma_prefetch_desc(&M.AT);
    tma_prefetch_desc(&M.B);
  }
  fence_proxy_async();
  __syncthreads();

  const long long per_tuple = 3LL * P.mtiles * P.ntiles;
  const long long nitems = per_tuple * P.ntuples;
  int stage = 0;
  uint32_t phase = 0;

  if (warp >= nwarps) {
    // ------------------------------------------------------------------ producer
    if (CREGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == nwarps && lane == 0) {
      const uint32_t tx = (uint32_t)((P.tu * P.tv + P.brows) * 128);
      for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tup = (int)(item / per_tuple);
        int rem = (int)(item - (long long)tup * per_tuple);
        const int cls = rem / (P.mtiles * P.ntiles);
        rem -= cls * P.mtiles * P.ntiles;
        const int mt = rem / P.ntiles, nt = rem - mt * P.ntiles;
        const TupleRec *rec = P.recs + tup;
        if (rec->fake) continue;  // FAKE_TUPLE (Tuples.hpp:43)
        // operands per class: piece 0 = A_x plain, piece 1 = A_x with (p,q) exchanged;
        //   class 0: A_a B_bc + A_b^T B_ac   class 1: A_a B_cb' + A_c^T B_ab   class 2: A_b B_ca' + A_c^T B_ba'
        // rec->by is stored in exactly this (class, piece) order
#pragma unroll
        for (int piece = 0; piece < 2; piece++) {
          int xslot = rec->ax[cls == 0 ? piece : (cls == 1 ? 2 * piece : 1 + piece)];
          int bslot = rec->by[2 * cls + piece];
          const CUtensorMap *tm, *tmb = &M.B;
          if (xslot >= P.ownedA) { xslot -= P.ownedA; tm = piece ? &M.ATc : &M.Ac; }
          else tm = piece ? &M.AT : &M.A;
          if (bslot >= P.ownedB) { bslot -= P.ownedB; tmb = &M.Bc; }
          const int u0 = (mt % P.utiles) * P.tu, v0 = (mt / P.utiles) * P.tv;
          for (int kc = 0; kc < P.nk; kc++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            unsigned char *sb = base + (size_t)stage * stage_bytes;
            mbar_expect_tx(&full_bar[stage], tx);
            tma_load_4d(sb, tm, &full_bar[stage], kc * KC, u0, v0, xslot);
            tma_load_3d(sb + (size_t)P.arows * 128, tmb, &full_bar[stage], kc * KC, nt * NI * 8, bslot);
            if (++stage == P.nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  if (CREGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CREGS ? CREGS : 24));
  const int g = lane >> 2, t = lane & 3;
  const int perm = 2 * (g & 3) + (g >> 2);  // tile row (mod 8) held by fragment row g
  const uint32_t offA = (uint32_t)((warp * MI * 8 + perm) * 128 + (t & 1) * 8);
  const uint32_t offB = (uint32_t)((P.arows + perm) * 128 + (t & 1) * 8);
  uint32_t cs[4];
#pragma unroll
  for (int s = 0; s < 4; s++) cs[s] = (uint32_t)(((2 * s + (t >> 1)) ^ perm) * 16);
  const int nchunks = 2 * P.nk;
  const int tile_rows = P.tu * P.tv;

  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tup = (int)(item / per_tuple);
    int rem = (int)(item - (long long)tup * per_tuple);
    const int cls = rem / (P.mtiles * P.ntiles);
    rem -= cls * P.mtiles * P.ntiles;
    const int mt = rem / P.ntiles, nt = rem - mt * P.ntiles;
    if (P.recs[tup].fake) continue;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int ch = 0; ch < nchunks; ch++) {
      mbar_wait(&full_bar[stage], phase);
      const unsigned char *sb = base + (size_t)stage * stage_bytes;
      // one k-step = 4 values of kappa: MI + NI fragment loads feed MI x NI DMMA.8x8x4
      auto kstep = [&](uint32_t col) {
        double af[MI], bf[NI];
#pragma unroll
        for (int i = 0; i < MI; i++) af[i] = *reinterpret_cast<const double *>(sb + offA + i * 1024 + col);
#pragma unroll
        for (int j = 0; j < NI; j++) bf[j] = *reinterpret_cast<const double *>(sb + offB + j * 1024 + col);
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
          for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      };
      if (ch != P.nk - 1 && ch != nchunks - 1) {
#pragma unroll
        for (int s = 0; s < 4; s++) kstep(cs[s]);
      } else {
        // last chunk of an operand pair: only the k-steps that hold data (the rest of the 16-wide
        // chunk is zero padding; last_steps = 4 when Kp has no padding)
        kstep(cs[0]);
        if (P.last_steps > 1) kstep(cs[1]);
        if (P.last_steps > 2) kstep(cs[2]);
        if (P.last_steps > 3) kstep(cs[3]);
      }
      // Hand the stage back only after every fragment load of it has delivered its data: the
      // producer's refill is a TMA write (async proxy) over rows this warp read through the
      // generic proxy.  The ordering construct is the proxy fence every lane executes before the
      // warp-level rendezvous: generic-proxy reads (the LDS above) -> fence.proxy.async ->
      // __syncwarp -> mbarrier.arrive (release) -> producer's wait (acquire) -> TMA write.  Without
      // it nothing in the source ties the arrive to the loads (it has no register dependency) and
      // ptxas once scheduled it between the last LDS and the DMMAs consuming them in a
      // straight-line variant of this loop body: run-to-run different cubes
      // (profiles/r01_ring_release_race.txt).  tools/check_sass_order.py (run by
      // tests/test_host.py) stays as a second guard: it asserts  last LDS < FENCE < WARPSYNC <
      // arrive  and  last DMMA < arrive  in the SASS of every instantiation.
      release_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == P.nstages) { stage = 0; phase ^= 1; }
    }

    // epilogue: C fragment (row g, cols 2t, 2t+1) -> tile row/col through the same permutation.
    // Every class cube is stored at Tijk's own (i,j,k) -- class 0 (u,v,n) = (i,j,k), class 1
    // (i,k,j), class 2 (j,k,i) -- in the 8x8x8-blocked layout kernel 2 streams (cube_offset):
    // the offset is separable, off(i,j,k) = gi(i) + gj(j) + gk(k).
    double *Rc = P.R + (size_t)tup * P.tuple_stride + (size_t)cls * P.cube_stride;
    const int u0 = (mt % P.utiles) * P.tu, v0 = (mt / P.utiles) * P.tv;
    const unsigned nb = (unsigned)(P.No + 7) >> 3;
    // tile stride / in-tile stride of the coordinate the column index n plays in this class
    const unsigned cT = cls == 0 ? nb * nb * 512u : (cls == 1 ? nb * 512u : 512u), cS = cls == 0 ? 64u : (cls == 1 ? 8u : 1u);
    // ... and of the row coordinates u (fast) and v
    const unsigned uT = cls == 2 ? nb * 512u : 512u, uS = cls == 2 ? 8u : 1u;
    const unsigned vT = cls == 0 ? nb * 512u : nb * nb * 512u, vS = cls == 0 ? 8u : 64u;
    unsigned cb[2];  // column offset of fragment column e at j = 0: col = nt NI 8 + j 8 + pc, pc < 8
    int pc[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int cf = 2 * t + e;
      pc[e] = 2 * (cf & 3) + (cf >> 2);
      cb[e] = (unsigned)(nt * NI) * cT + (unsigned)pc[e] * cS;
    }
    int ul = (warp * MI * 8 + perm) % P.tu, vl = (warp * MI * 8 + perm) / P.tu;
#pragma unroll
    for (int i = 0; i < MI; i++) {
      const int rl = warp * MI * 8 + i * 8 + perm;
      const int u = u0 + ul, v = v0 + vl;
      const unsigned m = (unsigned)(u >> 3) * uT + (unsigned)(u & 7) * uS + (unsigned)(v >> 3) * vT + (unsigned)(v & 7) * vS;
      ul += 8;
      while (ul >= P.tu) { ul -= P.tu; vl++; }
      if (rl < tile_rows && u < P.No && v < P.No) {
#pragma unroll
        for (int j = 0; j < NI; j++) {
#pragma unroll
          for (int e = 0; e < 2; e++)
            if (nt * NI * 8 + j * 8 + pc[e] < P.No) Rc[m + cb[e] + (unsigned)j * cT] = acc[i][j][e];
        }
      }
    }
  }
}

}  // namespace ab
