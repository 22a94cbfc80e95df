// The B200 device engine behind include/atrip_b200.h: context, HBM stores, tuple lists, the
// batch loop over the two hot kernels (contraction.cuh, reduction.cuh) and the C-ABI.
//
// There is deliberately no CPU path in this file: every compute entry point needs the CUDA
// device and fails loudly without it.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/atrip_b200.h"
#include "comm.hpp"
#include "common.cuh"
#include "contraction.cuh"
#include "reduction.cuh"
#include "reduction_async.cuh"
#include "reduction_z.cuh"
#include "schedule.hpp"
#include "stores.cuh"
#include "tuples.hpp"

using namespace ab;

namespace {

thread_local std::string g_error;

struct Fail {
  std::string msg;
};
#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw Fail{std::string("CUDA: ") + cudaGetErrorString(e__) + " in " #expr " (" __FILE__ ":" + \
                 std::to_string(__LINE__) + ")"};                                                  \
  } while (0)
#define REQUIRE(cond, text)                 \
  do {                                      \
    if (!(cond)) throw Fail{std::string(text)}; \
  } while (0)

// ------------------------------------------------------------------ contraction kernel registry
struct KernelVariant {
  int MI, NI, maxt, cregs;
  const void *fn;
};
#define VARIANT(mi, ni, maxt) KernelVariant{mi, ni, maxt, 0, (const void *)contract_kernel<mi, ni, maxt>}
// 8 (or 4) consumer warps + a producer warpgroup, registers re-allocated between the roles with
// setmaxnreg (contraction.cuh): 232 registers per consumer thread, no spills in any of them
#define VARIANT_R(mi, ni) KernelVariant{mi, ni, 384, 232, (const void *)contract_kernel<mi, ni, 384, 232>}
// accumulators take 4*MI*NI registers; the thread cap follows from the 64K register file (uniform
// allocation: 16K registers per SM sub-partition, e.g. 3 warps of a 9-warp block on one of them -> 168)
const KernelVariant kVariants[] = {
    VARIANT_R(5, 5), VARIANT_R(4, 6), VARIANT_R(5, 6), VARIANT_R(4, 7), VARIANT_R(5, 7), VARIANT_R(4, 8),
    VARIANT_R(5, 8), VARIANT_R(3, 10), VARIANT_R(4, 10), VARIANT_R(2, 13), VARIANT_R(3, 13),
    VARIANT(2, 1, 512), VARIANT(3, 1, 512), VARIANT(4, 1, 512), VARIANT(5, 1, 512),
    VARIANT(2, 2, 512), VARIANT(3, 2, 512), VARIANT(4, 2, 512), VARIANT(5, 2, 512),
    VARIANT(2, 3, 512), VARIANT(3, 3, 512), VARIANT(4, 3, 512), VARIANT(5, 3, 512),
    VARIANT(2, 4, 512), VARIANT(3, 4, 512), VARIANT(4, 4, 512), VARIANT(5, 4, 416),
    VARIANT(2, 5, 512), VARIANT(3, 5, 512), VARIANT(4, 5, 416), VARIANT(5, 5, 288),
    VARIANT(2, 6, 512), VARIANT(3, 6, 416), VARIANT(4, 6, 288), VARIANT(5, 6, 288),
    VARIANT(2, 7, 512), VARIANT(3, 7, 416), VARIANT(4, 7, 288),
    VARIANT(2, 8, 512), VARIANT(3, 8, 416), VARIANT(4, 8, 288),
    VARIANT(2, 10, 416), VARIANT(3, 10, 288),
    VARIANT(2, 13, 288),
};

struct ContractPlan {
  const KernelVariant *k = nullptr;
  int nw = 0, tu = 0, tv = 0, utiles = 0, mtiles = 0, ntiles = 0, arows = 0, brows = 0, nstages = 0;
  size_t smem = 0;
  double useful = 0;
};

// Pick the kernel variant, consumer-warp count and row tile (tu x tv) for this No.
//   useful   = fraction of issued DMMA work that lands inside the No^2 x No class matrix
//   smsp_eff = consumer warps are dealt round-robin to the 4 SM sub-partitions, each with its own
//              tensor pipe: NW % 4 != 0 leaves pipes idle (ncu r01: NW = 10 -> 81.6 % DMMA active)
//   warp_eff = tools/fp64_peak.cu: one warp per sub-partition reaches ~89 % of the DMMA ceiling,
//              two or more reach it
ContractPlan plan_contraction(int No, size_t smem_limit) {
  ContractPlan best;
  double best_score = -1;
  const char *env = std::getenv("ATRIP_B200_NO_SETMAXNREG");  // developer knob: A/B against the uniform allocation
  const bool no_r = env && std::atoi(env) != 0;
  for (const auto &k : kVariants) {
    if (no_r && k.cregs) continue;
    const int ntiles = (No + k.NI * 8 - 1) / (k.NI * 8);
    for (int nw = 4; nw <= k.maxt / 32 - (k.cregs ? 4 : 1); nw++) {
      if (k.cregs && nw % 4) continue;  // setmaxnreg works on whole warpgroups
      const int arows = nw * k.MI * 8;
      const size_t stage = contract_stage_bytes(arows, k.NI);
      const int nstages = (int)std::min<size_t>(MAX_STAGES, (smem_limit - 2048) / stage);
      if (nstages < 3) continue;
      const double smsp_eff = (double)nw / (4.0 * ((nw + 3) / 4));
      const double warp_eff = nw >= 8 ? 1.0 : 0.89;
      const double stage_eff = nstages >= 4 ? 1.0 : 0.97;
      // fewer LDS per DMMA with larger fragment grids (smem bandwidth headroom)
      const double frag_eff = 1.0 - 0.02 * (k.MI + k.NI) / (double)(k.MI * k.NI);
      for (int tu = std::min(std::min(No, 256), arows); tu >= 1; tu--) {  // ties: prefer long u runs
        const int tv = std::min(std::min(No, 256), arows / tu);
        const int utiles = (No + tu - 1) / tu, vtiles = (No + tv - 1) / tv;
        const double useful =
            (double)No * No * No / ((double)utiles * vtiles * arows * ntiles * k.NI * 8);
        const double score = useful * smsp_eff * warp_eff * stage_eff * frag_eff;
        if (score > best_score + 1e-9) {
          best_score = score;
          best.k = &k;
          best.nw = nw;
          best.tu = tu;
          best.tv = tv;
          best.utiles = utiles;
          best.mtiles = utiles * vtiles;
          best.ntiles = ntiles;
          best.arows = arows;
          best.brows = std::min(k.NI * 8, No);
          best.nstages = nstages;
          best.smem = stage * nstages + 1024;
          best.useful = useful;
        }
      }
    }
  }
  return best;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    REQUIRE(p && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

void make_map(CUtensorMap *tm, void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
              const uint32_t *box) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, base, gdim, gstr, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Fail{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r)};
}

template <typename T>
T *dalloc(size_t n) {
  T *p = nullptr;
  CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}

// scratch device buffer of one call: freed on every exit path, including a thrown Fail
template <typename T>
struct Scratch {
  T *p = nullptr;
  Scratch() = default;
  explicit Scratch(size_t n) : p(dalloc<T>(n)) {}
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  void alloc(size_t n) { p = dalloc<T>(n); }
  ~Scratch() {
    if (p) cudaFree(p);
  }
};

}  // namespace

constexpr int REC_RING = 4;  // batches the host may run ahead of the device
constexpr int NXS = 4;       // copy streams of the P2P transport (the pulls of a batch are dealt round-robin)
constexpr int TRING = 8;     // per-batch timing events are harvested TRING batches later

// timing events of one batch: [w0, c0] = gap in front of the contraction launch (slice fetch not yet
// landed, cube buffer not yet reduced), [c0, c1] contraction, [r0, r1] reduction + batch sum
struct BatchTimers {
  cudaEvent_t w0 = nullptr, c0 = nullptr, c1 = nullptr, r0 = nullptr, r1 = nullptr;
};

struct atrip_b200_ctx {
  atrip_b200_config cfg{};
  int No = 0, Nv = 0, Kp = 0;
  int cplx = 0;   // field: 0 FP64 real, 1 std::complex<double> (stores.cuh "complex field")
  int Klen = 0;   // valid contraction length: No + Nv (real), 2 (No + Nv) (complex); Kp = Klen padded to 16
  int nsm = 0;
  size_t smem_limit = 0;
  // contraction (high priority); reduction of the previous batch (low priority, runs beside the
  // next contraction on the same SMs); slice exchange (side stream)
  cudaStream_t stream = nullptr, rstream = nullptr, xstream = nullptr;
  cudaStream_t xs[NXS]{};  // xs[0] == xstream
  cudaStream_t xv = nullptr;  // P2P pulls of Vabij blocks: only the reduction reads them (own ordering, pull_step)
  cudaEvent_t xvdone[4]{};
  cudaEvent_t xfork = nullptr, xjoin[NXS]{};
  cudaEvent_t ev[6]{};
  BatchTimers bt[TRING];

  // stores: owned slices in the layouts of stores.cuh, slot numbering of schedule.hpp
  ShardMap map;                // storage sharding (replica: n = 1)
  int64_t owned[3] = {0, 0, 0};  // owned slots per kind (KA, KB, KV)
  double *AX = nullptr, *BY = nullptr, *VIJ = nullptr;
  double *AXJ = nullptr, *BYJ = nullptr;  // (cT): J tensors in the same layouts
  double *eps_i = nullptr, *eps_a = nullptr, *Tai = nullptr;
  int *xtab = nullptr, *btab = nullptr, *vtab = nullptr;  // global id -> owned slot or -1 (ingest)
  int *xlist = nullptr, *ylist = nullptr, *zlist = nullptr, *tflag = nullptr, *vy = nullptr, *vz = nullptr;
  bool have_J = false;

  // fetch caches (sharded stores only): cap[kind] slots per store, managed by SliceCache
  // (schedule.hpp): slices stay until their slot is re-assigned, two batches after their last use
  int64_t cap[3] = {0, 0, 0};
  double *cA = nullptr, *cB = nullptr, *cV = nullptr, *cAJ = nullptr, *cBJ = nullptr;
  SliceCache cache;
  int64_t serial = 0;  // number of the next batch (monotonic over runs)

  // tuples
  std::vector<Tuple> tuples;

  // work buffers
  int batch = 0;
  double *R[2] = {nullptr, nullptr}, *RJ[2] = {nullptr, nullptr};  // class cubes, double-buffered by batch parity
  double *e_tuple = nullptr, *d_total = nullptr;
  TupleRec *h_recs = nullptr, *d_recs = nullptr;  // [REC_RING][batch]
  cudaEvent_t rec_ev[REC_RING]{};
  uint64_t rec_uses = 0;

  // kernel plan
  ContractPlan plan;
  ContractMaps maps, mapsJ;    // real field / complex variant 0 ([Re A | -Im A] rows of the AX slices)
  ContractMaps maps1, mapsJ1;  // complex variant 1 ([Im A | Re A] rows)

  // slice exchange
  ncclComm_t comm = nullptr;
  size_t req_cap = 0;                                     // ints per request list
  int32_t *h_req_send = nullptr, *h_req_recv = nullptr;   // pinned [2][nranks][req_cap]
  int32_t *d_req_send = nullptr, *d_req_recv = nullptr;
  cudaEvent_t xdone[4]{}, cdone[4]{};
  cudaEvent_t evV[4]{}, evJ[4]{}, rdone[4]{};               // contraction (V / J pass) and reduction of batch k done
  double *d_reduce = nullptr;                             // all-reduce scratch
  int transport = 0;                                      // 1 NCCL send/recv, 2 P2P pulls (copy engines)
  int comm_sms = 0;                                       // SMs the contraction leaves to NCCL kernels
  std::vector<double *> peer[5];                          // P2P: peers' AX, BY, VIJ, AXJ, BYJ (IPC mapped)
  bool stores_dirty = true;                               // filled/loaded since the last rank barrier
  double exch_bytes = 0, exch_msgs = 0;                   // of the last run (received)

  // staging for ingest
  double *h_stage[2] = {nullptr, nullptr}, *d_stage[2] = {nullptr, nullptr};
  size_t stage_elems = 0;
  cudaEvent_t stage_ev[2]{};

  double timing[16] = {0};
  int last_nt = 0, last_buf = 0;  // tuples and cube buffer of the last batch run (debug checksum)
  bool reduce_async = false;      // bulk-copy reduction kernel (reduction_async.cuh): default for the real field
  bool reduce_reverse = false;    // ATRIP_B200_REDUCE=async-rev: ... walking the batch last tuple first
  bool solo = false;              // ATRIP_B200_SOLO_SHARD=1 (profiling only): one rank of a sharded job runs alone, on
                                  // tuples whose slices it owns itself (e.g. the c4 kernel shapes on one GPU)
};

namespace {

StoreDims dims_of(const atrip_b200_ctx *c) { return StoreDims{c->No, c->Nv, c->Kp, c->cplx}; }

// doubles per slice.  Complex field: an AX slice holds both variants, a VIJ slice is interleaved
size_t slice_elems(const atrip_b200_ctx *c, int kind) {
  const size_t No = c->No, Kp = c->Kp, z = c->cplx ? 2 : 1;
  return kind == KA ? z * No * No * Kp : (kind == KB ? No * Kp : z * No * No);
}
size_t esz(const atrip_b200_ctx *c) { return c->cplx ? 2 : 1; }        // doubles per tensor element
int ncubes(const atrip_b200_ctx *c) { return c->cplx ? 6 : 3; }        // class cubes per tuple

// variant: which half of a complex AX slice the A maps address (0 for the real field)
void build_maps(atrip_b200_ctx *c, double *AX, uint64_t nA, double *BY, uint64_t nB, CUtensorMap *tA,
                CUtensorMap *tAT, CUtensorMap *tB, int variant = 0) {
  const uint64_t No = c->No, Kp = c->Kp;
  {
    const uint64_t slot = slice_elems(c, KA) * 8;  // bytes between slices (complex: two variants each)
    const uint64_t dims[4] = {Kp, No, No, std::max<uint64_t>(nA, 1)};
    const uint64_t strP[3] = {Kp * 8, No * Kp * 8, slot};
    const uint64_t strT[3] = {No * Kp * 8, Kp * 8, slot};
    const uint32_t box[4] = {KC, (uint32_t)c->plan.tu, (uint32_t)c->plan.tv, 1};
    double *base = AX + (size_t)variant * No * No * Kp;
    make_map(tA, base, 4, dims, strP, box);
    make_map(tAT, base, 4, dims, strT, box);
  }
  {
    const uint64_t dims[3] = {Kp, No, std::max<uint64_t>(nB, 1)};
    const uint64_t str[2] = {Kp * 8, No * Kp * 8};
    const uint32_t box[3] = {KC, (uint32_t)c->plan.brows, 1};
    make_map(tB, BY, 3, dims, str, box);
  }
}

// (re)build all tensor maps: owned stores, and the fetch caches when there are any
void build_all_maps(atrip_b200_ctx *c) {
  for (int var = 0; var <= c->cplx; var++) {
    ContractMaps &M = var ? c->maps1 : c->maps, &MJ = var ? c->mapsJ1 : c->mapsJ;
    build_maps(c, c->AX, c->owned[KA], c->BY, c->owned[KB], &M.A, &M.AT, &M.B, var);
    if (c->cA) build_maps(c, c->cA, c->cap[KA], c->cB, c->cap[KB], &M.Ac, &M.ATc, &M.Bc, var);
    else { M.Ac = M.A; M.ATc = M.AT; M.Bc = M.B; }
    if (c->cfg.with_J) {
      build_maps(c, c->AXJ, c->owned[KA], c->BYJ, c->owned[KB], &MJ.A, &MJ.AT, &MJ.B, var);
      if (c->cAJ) build_maps(c, c->cAJ, c->cap[KA], c->cBJ, c->cap[KB], &MJ.Ac, &MJ.ATc, &MJ.Bc, var);
      else { MJ.Ac = MJ.A; MJ.ATc = MJ.AT; MJ.Bc = MJ.B; }
    }
  }
}

bool sharded(const atrip_b200_ctx *c) { return c->map.n > 1; }

// fetch caches sized for the current tuple list (schedule.hpp: cache_need = distinct remote slices
// of any window of one batch): three windows -- the batch computing, the batch being fetched and the
// batch prefetched for the next run call; grown on demand
void ensure_caches(atrip_b200_ctx *c, const int64_t need[3]) {
  if (!sharded(c)) return;
  // a debug tuple (12 slices, all remote in the worst case) must always fit
  // ... and more while it is cheap (<= 4 GiB): slices of the slowly varying indices then survive
  // until their run of the tuple list comes back (c2 at 8 ranks: 187 -> 112 MB fetched per batch)
  int64_t mult = 3;
  auto bytes_of = [&](int64_t m) {
    double b = 0;
    for (int k = 0; k < 3; k++) b += (double)m * need[k] * slice_elems(c, k) * 8 * ((c->cfg.with_J && k != KV) ? 2 : 1);
    return b;
  };
  while (mult < 8 && bytes_of(mult + 1) <= 4.0 * (1u << 30)) mult++;
  const int64_t want[3] = {std::max<int64_t>(mult * need[KA], 3), std::max<int64_t>(mult * need[KB], 6),
                           std::max<int64_t>(mult * need[KV], 3)};
  if (c->cA && want[KA] <= c->cap[KA] && want[KB] <= c->cap[KB] && want[KV] <= c->cap[KV]) return;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaStreamSynchronize(c->xstream));
  for (double **p : {&c->cA, &c->cB, &c->cV, &c->cAJ, &c->cBJ})
    if (*p) { cudaFree(*p); *p = nullptr; }
  for (int k = 0; k < 3; k++) c->cap[k] = std::max(c->cap[k], want[k]);
  REQUIRE(c->owned[KB] + c->cap[KB] < (1LL << 31) && c->owned[KV] + c->cap[KV] < (1LL << 31),
          "too many slots for 32-bit slot numbers");
  c->cA = dalloc<double>(c->cap[KA] * slice_elems(c, KA));
  c->cB = dalloc<double>(c->cap[KB] * slice_elems(c, KB));
  c->cV = dalloc<double>(c->cap[KV] * slice_elems(c, KV));
  if (c->cfg.with_J) {
    c->cAJ = dalloc<double>(c->cap[KA] * slice_elems(c, KA));
    c->cBJ = dalloc<double>(c->cap[KB] * slice_elems(c, KB));
  }
  c->cache.reset(c->cap);
  build_all_maps(c);
}

void ensure_stage(atrip_b200_ctx *c, size_t elems) {
  if (c->stage_elems >= elems) return;
  for (int i = 0; i < 2; i++) {
    if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    CUDA_OK(cudaMallocHost(&c->h_stage[i], elems * sizeof(double)));
    c->d_stage[i] = dalloc<double>(elems);
    if (!c->stage_ev[i]) CUDA_OK(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
  }
  c->stage_elems = elems;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// stream a host tensor through the double-buffered staging pair, one chunk at a time;
// `consume(chunk_index, device_ptr)` enqueues the re-tiling kernel on c->stream
template <typename F>
void stream_chunks(atrip_b200_ctx *c, const double *host, size_t nchunks, size_t chunk_elems, F consume) {
  ensure_stage(c, chunk_elems);
  const bool pinned = is_pinned(host);
  for (size_t k = 0; k < nchunks; k++) {
    const int s = (int)(k & 1);
    // the kernel that last read d_stage[s] (and the copy out of h_stage[s]) must be done
    CUDA_OK(cudaEventSynchronize(c->stage_ev[s]));
    const double *src = host + k * chunk_elems;
    if (!pinned) {
      std::memcpy(c->h_stage[s], src, chunk_elems * sizeof(double));
      src = c->h_stage[s];
    }
    CUDA_OK(cudaMemcpyAsync(c->d_stage[s], src, chunk_elems * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    consume(k, c->d_stage[s]);
    CUDA_OK(cudaEventRecord(c->stage_ev[s], c->stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

const void *contract_fn(const atrip_b200_ctx *c) { return c->plan.k->fn; }

// avar: AX variant read by this launch (complex field: 0 -> Re cubes, 1 -> Im cubes)
void launch_contract(atrip_b200_ctx *c, const TupleRec *d_recs, int ntuples, bool useJ, int buf, int avar = 0) {
  ContractParams P;
  P.No = c->No;
  P.Nv = c->Nv;
  P.Kp = c->Kp;
  P.nk = c->Kp / KC;
  // k-steps of the last chunk that hold data (the kernel skips the zero padding behind them)
  P.last_steps = std::max(1, std::min(4, (c->Klen - (P.nk - 1) * KC + 3) / 4));
  P.tu = c->plan.tu;
  P.tv = c->plan.tv;
  P.utiles = c->plan.utiles;
  P.mtiles = c->plan.mtiles;
  P.ntiles = c->plan.ntiles;
  P.arows = c->plan.arows;
  P.brows = c->plan.brows;
  P.nstages = c->plan.nstages;
  P.ntuples = ntuples;
  P.ownedA = (int)c->owned[KA];
  P.ownedB = (int)c->owned[KB];
  P.recs = d_recs;
  P.cube_stride = cube_blocked_elems(c->No);
  P.tuple_stride = ncubes(c) * P.cube_stride;
  P.R = (useJ ? c->RJ[buf] : c->R[buf]) + (size_t)avar * 3 * P.cube_stride;
  const long long nitems = 3LL * P.mtiles * P.ntiles * ntuples;
  // NCCL transport: leave a few SMs to the send/recv kernels of the side stream, otherwise they
  // only run in the gaps between two persistent contraction launches
  const int grid = (int)std::min<long long>(c->nsm - c->comm_sms, nitems);
  if (grid <= 0) return;
  ContractMaps *maps = useJ ? (avar ? &c->mapsJ1 : &c->mapsJ) : (avar ? &c->maps1 : &c->maps);
  void *args[2] = {(void *)maps, (void *)&P};
  CUDA_OK(cudaLaunchKernel(contract_fn(c), dim3(grid), dim3((c->plan.nw + (c->plan.k->cregs ? 4 : 1)) * 32), args, c->plan.smem, c->stream));
}

ReduceParams reduce_params(atrip_b200_ctx *c, const TupleRec *d_recs, int ntuples, bool ct, int buf) {
  ReduceParams P;
  P.No = c->No;
  P.Nv = c->Nv;
  P.ntuples = ntuples;
  P.recs = d_recs;
  P.R = ct ? c->RJ[buf] : c->R[buf];
  P.RZ = c->R[buf];
  P.cube_stride = cube_blocked_elems(c->No);
  P.eps_i = c->eps_i;
  P.eps_a = c->eps_a;
  P.Tai = c->Tai;
  P.VIJ = c->VIJ;
  P.VIJc = c->cV ? c->cV : c->VIJ;
  P.ownedV = (int)c->owned[KV];
  P.e_tuple = c->e_tuple;
  // CTAs per tuple: fill the 2-CTA/SM slots in whole waves (a batch has few tuples when No is
  // large), but keep several orbits per CTA so its prologue (eps, Tai rows) stays amortised
  const int nb = (c->No + RT - 1) / RT, orbits = nb * (nb + 1) * (nb + 2) / 6;
  // resident CTAs of 128 threads per SM: 2 of the bulk-copy kernel (shared memory), 4 of the others
  const double slots = ((c->reduce_async && !ct && !c->cplx) ? 2.0 : 4.0) * c->nsm;
  int best = 1;
  double best_score = -1;
  for (int ns = 1; ns <= std::min(orbits, 64); ns++) {
    const double waves = (double)ntuples * ns / slots;
    const double fill = waves / std::ceil(waves);
    const double per_cta = (double)orbits / ns;
    const double score = fill * (per_cta / (per_cta + 1.0)) * (waves >= 2.5 ? 1.0 : 0.8 + 0.08 * waves);
    if (score > best_score + 1e-9) { best_score = score; best = ns; }
  }
  P.nsplit = best;
  P.reverse = c->reduce_reverse ? 1 : 0;
  if (const char *e = std::getenv("ATRIP_B200_NSPLIT")) {  // developer knob: force the orbit split
    const int v = std::atoi(e);
    if (v >= 1) P.nsplit = std::min(v, std::min(orbits, 64));
  }
  return P;
}

void launch_reduce(atrip_b200_ctx *c, const TupleRec *d_recs, int ntuples, bool ct, int buf, double *total) {
  if (ntuples <= 0) return;
  ReduceParams P = reduce_params(c, d_recs, ntuples, ct, buf);
  const size_t smem = reduce_smem_bytes(c->No, ct);
  const dim3 grid(ntuples, P.nsplit);
  if (c->cplx) {
    const size_t smz = reduce_z_smem_bytes(c->No, ct);
    if (ct) reduce_z_kernel<true><<<grid, REDUCE_THREADS, smz, c->rstream>>>(P);
    else reduce_z_kernel<false><<<grid, REDUCE_THREADS, smz, c->rstream>>>(P);
  } else if (c->reduce_async && !ct) {
    reduce_async_kernel<<<grid, RA_THREADS, reduce_async_smem_bytes(c->No), c->rstream>>>(P);
  } else if (ct) reduce_kernel<true><<<grid, REDUCE_THREADS, smem, c->rstream>>>(P);
  else reduce_kernel<false><<<grid, REDUCE_THREADS, smem, c->rstream>>>(P);
  CUDA_OK(cudaGetLastError());
  accumulate_kernel<<<1, 256, 0, c->rstream>>>(c->e_tuple, ntuples * P.nsplit, total);
  CUDA_OK(cudaGetLastError());
}

template <typename T>
int *upload_ints(const std::vector<T> &h) {
  std::vector<int> v(h.begin(), h.end());
  int *d = dalloc<int>(v.size());
  CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return d;
}

void create_impl(atrip_b200_ctx *c) {
  const auto &cfg = c->cfg;
  REQUIRE(cfg.No >= 1 && cfg.Nv >= 1, "No and Nv must be positive");
  REQUIRE(cfg.No <= 256, "No > 256 is not supported by the TMA box of the contraction kernel");
  REQUIRE(cfg.nranks >= 1 && cfg.rank >= 0 && cfg.rank < cfg.nranks, "bad rank / nranks");
  REQUIRE(cfg.Nv * cfg.Nv + cfg.Nv < (1LL << 31), "Nv too large for 32-bit pair tables");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Fail{std::string("no usable CUDA device (") + cudaGetErrorString(e) +
               "); the atrip_b200 engine has no CPU fallback"};
  REQUIRE(cfg.device >= 0 && cfg.device < ndev, "device ordinal out of range");
  CUDA_OK(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, cfg.device));
  REQUIRE(prop.major >= 10, "atrip_b200 needs an sm_100a device (B200)");
  c->nsm = prop.multiProcessorCount;
  c->smem_limit = prop.sharedMemPerBlockOptin;
  c->No = (int)cfg.No;
  c->Nv = (int)cfg.Nv;
  REQUIRE(cfg.field == 0 || cfg.field == 1, "field must be 0 (FP64 real) or 1 (FP64 complex)");
  c->cplx = cfg.field;
  c->Klen = (int)((cfg.No + cfg.Nv) * (c->cplx ? 2 : 1));
  c->Kp = (c->Klen + KC - 1) / KC * KC;
  if (const char *e = std::getenv("ATRIP_B200_KPAD"))  // developer knob: extra all-zero K chunks
    c->Kp += KC * std::max(0, std::atoi(e));
  int prio_least = 0, prio_greatest = 0;
  CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_OK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
  CUDA_OK(cudaStreamCreateWithPriority(&c->rstream, cudaStreamNonBlocking, prio_least));
  CUDA_OK(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
  c->xs[0] = c->xstream;
  for (int i = 1; i < NXS; i++) CUDA_OK(cudaStreamCreateWithFlags(&c->xs[i], cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&c->xv, cudaStreamNonBlocking));
  for (auto &ev : c->xvdone) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->xfork, cudaEventDisableTiming));
  for (int i = 1; i < NXS; i++) CUDA_OK(cudaEventCreateWithFlags(&c->xjoin[i], cudaEventDisableTiming));
  for (auto &b : c->bt)
    for (cudaEvent_t *e : {&b.w0, &b.c0, &b.c1, &b.r0, &b.r1}) CUDA_OK(cudaEventCreate(e));
  for (auto &ev : c->ev) CUDA_OK(cudaEventCreate(&ev));
  for (auto &ev : c->rec_ev) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->xdone) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->cdone) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->evV) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->evJ) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->rdone) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

  c->plan = plan_contraction(c->No, c->smem_limit);
  REQUIRE(c->plan.k, "no contraction kernel variant fits this No");
  // The dynamic shared-memory cap is an attribute of the FUNCTION, shared by every context of the process:
  // a second engine with a smaller No must not lower it under a live engine (bench.py keeps its engine
  // while the golden case runs on a small one).  So every kernel gets the cap of the largest supported
  // problem (No = 256), not of this engine's No.
  const int kMaxNo = 256;
  auto set_cap = [&](const void *fn, size_t bytes) {  // dynamic cap <= opt-in limit - the kernel's static shared memory
    cudaFuncAttributes fa;
    CUDA_OK(cudaFuncGetAttributes(&fa, fn));
    const size_t room = c->smem_limit - std::min(c->smem_limit, fa.sharedSizeBytes);
    CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min(bytes, room)));
  };
  set_cap(contract_fn(c), c->smem_limit);
  set_cap((const void *)reduce_kernel<false>, reduce_smem_bytes(kMaxNo, false));
  set_cap((const void *)reduce_kernel<true>, reduce_smem_bytes(kMaxNo, true));
  if (const char *e = std::getenv("ATRIP_B200_SOLO_SHARD")) c->solo = std::atoi(e) != 0;
  // reduction kernel of the (T) pass, real field: the bulk-copy kernel (reduction_async.cuh; measured r02c:
  // 263 us vs 349 us per c2 launch).  ATRIP_B200_REDUCE=sync selects the register-staged kernel, =async-rev
  // the bulk-copy kernel walking the batch last tuple first (developer A/B knobs)
  c->reduce_async = !c->cplx;
  if (const char *e = std::getenv("ATRIP_B200_REDUCE")) {
    c->reduce_reverse = std::string(e) == "async-rev" && !c->cplx;
    if (std::string(e) == "sync") c->reduce_async = false;
  }
  set_cap((const void *)reduce_async_kernel, reduce_async_smem_bytes(kMaxNo));
  REQUIRE(reduce_async_smem_bytes(c->No) <= c->smem_limit && reduce_smem_bytes(c->No, true) <= c->smem_limit,
          "No too large for the reduction kernels");
  if (c->cplx) {
    REQUIRE(reduce_z_smem_bytes(c->No, true) <= c->smem_limit, "No too large for the complex reduction kernel");
    set_cap((const void *)reduce_z_kernel<false>, reduce_z_smem_bytes(kMaxNo, false));
    set_cap((const void *)reduce_z_kernel<true>, reduce_z_smem_bytes(kMaxNo, true));
  }

  // ---- which slices live here: everything (replica) or the slices this rank owns
  const int sn = (cfg.resident || cfg.nranks == 1) ? 1 : cfg.nranks;
  REQUIRE(cfg.transport >= 0 && cfg.transport <= 2, "transport must be 0 (default), 1 (NCCL) or 2 (P2P)");
  c->transport = sn == 1 ? 0 : (cfg.transport == 0 ? 2 : cfg.transport);
  if (c->transport == 1) {
    const char *e = std::getenv("ATRIP_B200_COMM_SMS");
    c->comm_sms = std::max(0, std::min(c->nsm / 4, e ? std::atoi(e) : 4));
  }
  c->map = ShardMap(cfg.Nv, sn, sn == 1 ? 0 : cfg.rank);
  const ShardMap &m = c->map;
  for (int k = 0; k < 3; k++) c->owned[k] = m.owned(k, m.me);
  REQUIRE(c->owned[KB] < (1LL << 31) && c->owned[KV] < (1LL << 31), "too many slots for 32-bit slot numbers");
  const int64_t Nv = c->Nv, NvNv = Nv * Nv;
  {
    std::vector<int> xtab((size_t)Nv, -1), xlist((size_t)c->owned[KA], 0);
    for (int64_t x = 0; x < Nv; x++)
      if (m.ownerA(x) == m.me) {
        xtab[(size_t)x] = (int)m.slotA(x);
        xlist[(size_t)m.slotA(x)] = (int)x;
      }
    std::vector<int> btab((size_t)(NvNv + Nv), -1), yl((size_t)c->owned[KB], 0), zl((size_t)c->owned[KB], 0),
        tf((size_t)c->owned[KB], 0);
    for (int64_t id = 0; id < NvNv + Nv; id++)
      if (m.ownerB(id) == m.me) {
        const int64_t s = m.slotB(id);
        btab[(size_t)id] = (int)s;
        yl[(size_t)s] = (int)(id < NvNv ? id % Nv : id - NvNv);
        zl[(size_t)s] = (int)(id < NvNv ? id / Nv : id - NvNv);
        tf[(size_t)s] = id >= NvNv;
      }
    std::vector<int> vtab((size_t)NvNv, -1), vy((size_t)c->owned[KV], 0), vz((size_t)c->owned[KV], 0);
    for (int64_t z = 0; z < Nv; z++)
      for (int64_t y = 0; y <= z; y++) {
        const int64_t s = m.localV(y, z);
        if (s < 0) continue;
        vtab[(size_t)(y + z * Nv)] = (int)s;
        vy[(size_t)s] = (int)y;
        vz[(size_t)s] = (int)z;
      }
    c->xtab = upload_ints(xtab);
    c->xlist = upload_ints(xlist);
    c->btab = upload_ints(btab);
    c->ylist = upload_ints(yl);
    c->zlist = upload_ints(zl);
    c->tflag = upload_ints(tf);
    c->vtab = upload_ints(vtab);
    c->vy = upload_ints(vy);
    c->vz = upload_ints(vz);
  }

  // ---- stores
  const size_t No = c->No;
  const size_t axn = c->owned[KA] * slice_elems(c, KA), byn = c->owned[KB] * slice_elems(c, KB),
               vn = c->owned[KV] * slice_elems(c, KV);
  c->AX = dalloc<double>(axn);
  c->BY = dalloc<double>(byn);
  c->VIJ = dalloc<double>(vn);
  CUDA_OK(cudaMemsetAsync(c->AX, 0, axn * 8, c->stream));
  CUDA_OK(cudaMemsetAsync(c->BY, 0, byn * 8, c->stream));
  CUDA_OK(cudaMemsetAsync(c->VIJ, 0, vn * 8, c->stream));
  if (cfg.with_J) {
    c->AXJ = dalloc<double>(axn);
    c->BYJ = dalloc<double>(byn);
    CUDA_OK(cudaMemsetAsync(c->AXJ, 0, axn * 8, c->stream));
    CUDA_OK(cudaMemsetAsync(c->BYJ, 0, byn * 8, c->stream));
  }
  c->eps_i = dalloc<double>(esz(c) * No);
  c->eps_a = dalloc<double>(esz(c) * Nv);
  c->Tai = dalloc<double>(esz(c) * No * Nv);

  // ---- work buffers: a batch keeps roughly 8 work items per SM in flight and <= 1 GiB of cubes
  const size_t cube3 = ncubes(c) * cube_blocked_elems(c->No);
  long long batch = cfg.batch_tuples;
  if (batch <= 0) {
    const long long per_tuple = 3LL * c->plan.mtiles * c->plan.ntiles;
    batch = std::max<long long>(c->nsm, (c->nsm * 24LL + per_tuple - 1) / per_tuple * 4);
    batch = std::min<long long>(batch, std::max<long long>(16, (1LL << 30) / (long long)(cube3 * 8)));
    batch = (batch + c->nsm - 1) / c->nsm * c->nsm;
    batch = std::min<long long>(batch, std::max<long long>(16, (1LL << 30) / (long long)(cube3 * 8)));
  }
  c->batch = (int)std::max<long long>(1, batch);
  for (int b = 0; b < 2; b++) {
    c->R[b] = dalloc<double>(cube3 * c->batch);
    if (cfg.with_J) c->RJ[b] = dalloc<double>(cube3 * c->batch);
  }
  c->e_tuple = dalloc<double>((size_t)c->batch * 64);
  c->d_total = dalloc<double>(2);
  c->d_reduce = dalloc<double>(16);
  CUDA_OK(cudaMallocHost(&c->h_recs, sizeof(TupleRec) * REC_RING * c->batch));
  c->d_recs = dalloc<TupleRec>((size_t)REC_RING * c->batch);

  if (sharded(c) && c->transport == 1) {
    c->req_cap = request_capacity_ints((size_t)c->batch);
    const size_t n = 2 * (size_t)cfg.nranks * c->req_cap;
    CUDA_OK(cudaMallocHost(&c->h_req_send, n * sizeof(int32_t)));
    CUDA_OK(cudaMallocHost(&c->h_req_recv, n * sizeof(int32_t)));
    c->d_req_send = dalloc<int32_t>(n);
    c->d_req_recv = dalloc<int32_t>(n);
  }
  if (sharded(c)) {
    const int64_t none[3] = {0, 0, 0};
    ensure_caches(c, none);
  }
  build_all_maps(c);
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

void destroy_impl(atrip_b200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->xstream) cudaStreamSynchronize(c->xstream);
  // P2P transport: a peer may still be pulling slices out of this rank's stores -- destroying a
  // sharded context is collective (like MPI_Finalize): wait until every rank got here
  if (c->comm && nccl().ok && c->transport == 2 && c->d_reduce && c->stream) {
    if (cudaMemsetAsync(c->d_reduce, 0, sizeof(double), c->stream) == cudaSuccess &&
        nccl().AllReduce(c->d_reduce, c->d_reduce, 1, ncclFloat64, ncclSum, c->comm, c->stream) == ncclSuccess)
      cudaStreamSynchronize(c->stream);
  }
  for (auto &v : c->peer)
    for (size_t p = 0; p < v.size(); p++)
      if (v[p] && (int)p != c->cfg.rank) cudaIpcCloseMemHandle(v[p]);
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  if (c->rstream) cudaStreamSynchronize(c->rstream);
  void *ptrs[] = {c->AX, c->BY, c->VIJ, c->AXJ, c->BYJ, c->eps_i, c->eps_a, c->Tai, c->xtab, c->btab,
                  c->vtab, c->xlist, c->ylist, c->zlist, c->tflag, c->vy, c->vz, c->R[0], c->R[1], c->RJ[0], c->RJ[1],
                  c->e_tuple, c->d_total, c->d_reduce, c->d_recs, c->d_stage[0], c->d_stage[1],
                  c->cA, c->cB, c->cV, c->cAJ, c->cBJ, c->d_req_send, c->d_req_recv};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto *h : c->h_stage)
    if (h) cudaFreeHost(h);
  for (void *h : {(void *)c->h_recs, (void *)c->h_req_send, (void *)c->h_req_recv})
    if (h) cudaFreeHost(h);
  auto kill = [](cudaEvent_t *evs, int n) {
    for (int i = 0; i < n; i++)
      if (evs[i]) cudaEventDestroy(evs[i]);
  };
  kill(c->ev, 6);
  for (auto &b : c->bt)
    for (cudaEvent_t *e : {&b.w0, &b.c0, &b.c1, &b.r0, &b.r1}) kill(e, 1);
  kill(&c->xfork, 1);
  kill(c->xjoin, NXS);
  kill(c->xvdone, 4);
  for (int i = 1; i < NXS; i++)
    if (c->xs[i]) cudaStreamDestroy(c->xs[i]);
  if (c->xv) cudaStreamDestroy(c->xv);
  kill(c->stage_ev, 2);
  kill(c->rec_ev, REC_RING);
  kill(c->xdone, 4);
  kill(c->cdone, 4);
  kill(c->evV, 4);
  kill(c->evJ, 4);
  kill(c->rdone, 4);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->rstream) cudaStreamDestroy(c->rstream);
  if (c->xstream) cudaStreamDestroy(c->xstream);
  delete c;
}

void rank_barrier(atrip_b200_ctx *c);

// The stores are about to change.  With the P2P transport the peers read this rank's stores
// directly: once the ranks have run tuples, a peer may still be pulling slices of its previous run,
// so (re)loading sharded stores after a run is collective -- wait until every rank got here.  The
// first run after the update synchronises the ranks again (run_list) before anybody reads.
void begin_store_update(atrip_b200_ctx *c) {
  if (sharded(c) && c->comm && c->transport == 2 && !c->stores_dirty) rank_barrier(c);
  c->stores_dirty = true;
}

int grid_for(size_t n, int nsm) { return (int)std::min<size_t>((n + 255) / 256, (size_t)nsm * 16); }

void fill_z_impl(atrip_b200_ctx *c, uint64_t seed, double scale) {
  begin_store_update(c);
  const StoreDims d = dims_of(c);
  const size_t No = c->No;
  const size_t nX = c->owned[KA], nB = c->owned[KB], nV = c->owned[KV];
  fill_small_z_kernel<<<grid_for(No * (size_t)c->Nv, c->nsm), 256, 0, c->stream>>>(
      c->eps_i, c->eps_a, c->Tai, d, synth_key(seed, T_EPS_I), synth_key(seed, T_EPS_A), synth_key(seed, T_TAI), scale);
  fill_AX_z_kernel<<<grid_for(nX * slice_elems(c, KA), c->nsm), 256, 0, c->stream>>>(
      c->AX, d, c->xlist, (int)nX, synth_key(seed, T_TABIJ), synth_key(seed, T_VIJKA), scale);
  fill_BY_z_kernel<<<grid_for(nB * slice_elems(c, KB), c->nsm), 256, 0, c->stream>>>(
      c->BY, d, c->ylist, c->zlist, c->tflag, nB, synth_key(seed, T_VABCI), synth_key(seed, T_TABIJ), scale);
  fill_VIJ_z_kernel<<<grid_for(nV * slice_elems(c, KV), c->nsm), 256, 0, c->stream>>>(c->VIJ, d, c->vy, c->vz, nV,
                                                                                     synth_key(seed, T_VABIJ), scale);
  if (c->cfg.with_J) {
    fill_AX_z_kernel<<<grid_for(nX * slice_elems(c, KA), c->nsm), 256, 0, c->stream>>>(
        c->AXJ, d, c->xlist, (int)nX, synth_key(seed, T_TABIJ), synth_key(seed, T_JIJKA), scale);
    fill_BY_z_kernel<<<grid_for(nB * slice_elems(c, KB), c->nsm), 256, 0, c->stream>>>(
        c->BYJ, d, c->ylist, c->zlist, c->tflag, nB, synth_key(seed, T_JABCI), synth_key(seed, T_TABIJ), scale);
    c->have_J = true;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->stores_dirty = true;
}

void fill_impl(atrip_b200_ctx *c, uint64_t seed, double scale) {
  if (c->cplx) return fill_z_impl(c, seed, scale);
  begin_store_update(c);
  const StoreDims d = dims_of(c);
  const size_t No = c->No;
  const size_t nX = c->owned[KA], nB = c->owned[KB], nV = c->owned[KV];
  fill_small_kernel<<<grid_for(No * (size_t)c->Nv, c->nsm), 256, 0, c->stream>>>(
      c->eps_i, c->eps_a, c->Tai, d, synth_key(seed, T_EPS_I), synth_key(seed, T_EPS_A), synth_key(seed, T_TAI), scale);
  fill_AX_kernel<<<grid_for(nX * slice_elems(c, KA), c->nsm), 256, 0, c->stream>>>(
      c->AX, d, c->xlist, (int)nX, synth_key(seed, T_TABIJ), synth_key(seed, T_VIJKA), scale);
  fill_BY_kernel<<<grid_for(nB * slice_elems(c, KB), c->nsm), 256, 0, c->stream>>>(
      c->BY, d, c->ylist, c->zlist, c->tflag, nB, synth_key(seed, T_VABCI), synth_key(seed, T_TABIJ), scale);
  fill_VIJ_kernel<<<grid_for(nV * slice_elems(c, KV), c->nsm), 256, 0, c->stream>>>(c->VIJ, d, c->vy, c->vz, nV,
                                                                                   synth_key(seed, T_VABIJ), scale);
  if (c->cfg.with_J) {
    fill_AX_kernel<<<grid_for(nX * slice_elems(c, KA), c->nsm), 256, 0, c->stream>>>(
        c->AXJ, d, c->xlist, (int)nX, synth_key(seed, T_TABIJ), synth_key(seed, T_JIJKA), scale);
    fill_BY_kernel<<<grid_for(nB * slice_elems(c, KB), c->nsm), 256, 0, c->stream>>>(
        c->BYJ, d, c->ylist, c->zlist, c->tflag, nB, synth_key(seed, T_JABCI), synth_key(seed, T_TABIJ), scale);
    c->have_J = true;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->stores_dirty = true;
}

void load_Tabij_impl(atrip_b200_ctx *c, const double *T) {
  const StoreDims d = dims_of(c);
  const size_t NvNv = (size_t)c->Nv * c->Nv;
  const dim3 grid((c->Nv + 31) / 32, (c->Nv + 31) / 32), block(32, 8);
  stream_chunks(c, T, (size_t)c->No * c->No, esz(c) * NvNv, [&](size_t k, const double *dev) {
    const int p = (int)(k % c->No), q = (int)(k / c->No);
    if (c->cplx) {
      const int g = grid_for(NvNv, c->nsm);
      ingest_Tabij_z_kernel<<<g, 256, 0, c->stream>>>(dev, d, p, q, c->AX, c->xtab, c->BY, c->btab);
      if (c->cfg.with_J) ingest_Tabij_z_kernel<<<g, 256, 0, c->stream>>>(dev, d, p, q, c->AXJ, c->xtab, c->BYJ, c->btab);
      return;
    }
    ingest_Tabij_kernel<<<grid, block, 0, c->stream>>>(dev, d, p, q, c->AX, c->xtab, c->BY, c->btab);
    if (c->cfg.with_J)
      ingest_Tabij_kernel<<<grid, block, 0, c->stream>>>(dev, d, p, q, c->AXJ, c->xtab, c->BYJ, c->btab);
  });
  CUDA_OK(cudaGetLastError());
}

void load_hhhp_impl(atrip_b200_ctx *c, const double *V, double *AX) {
  const StoreDims d = dims_of(c);
  const size_t cube = (size_t)c->No * c->No * c->No;
  const size_t z = esz(c);
  const size_t xper = std::max<size_t>(1, std::min<size_t>(c->Nv, (size_t)(32u << 20) / (z * cube * 8)));
  const size_t nchunks = (c->Nv + xper - 1) / xper;
  ensure_stage(c, z * xper * cube);
  auto ingest = [&](const double *dev, size_t x0, size_t nx) {
    if (c->cplx)
      ingest_Vijka_z_kernel<<<grid_for(nx * cube, c->nsm), 256, 0, c->stream>>>(dev, d, (int)x0, (int)nx, AX, c->xtab);
    else
      ingest_Vijka_kernel<<<grid_for(nx * cube, c->nsm), 256, 0, c->stream>>>(dev, d, (int)x0, (int)nx, AX, c->xtab);
  };
  // the last chunk may be shorter: stream_chunks copies full chunks, so handle the tail by hand
  const size_t full = c->Nv / xper;
  stream_chunks(c, V, full, z * xper * cube, [&](size_t k, const double *dev) { ingest(dev, k * xper, xper); });
  if (full < nchunks) {
    const size_t x0 = full * xper, nx = c->Nv - x0;
    CUDA_OK(cudaMemcpyAsync(c->d_stage[0], V + z * x0 * cube, z * nx * cube * 8, cudaMemcpyHostToDevice, c->stream));
    ingest(c->d_stage[0], x0, nx);
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  CUDA_OK(cudaGetLastError());
}

void load_Vabij_impl(atrip_b200_ctx *c, const double *V) {
  const StoreDims d = dims_of(c);
  const size_t NvNv = (size_t)c->Nv * c->Nv;
  stream_chunks(c, V, (size_t)c->No * c->No, esz(c) * NvNv, [&](size_t k, const double *dev) {
    if (c->cplx)
      ingest_Vabij_z_kernel<<<grid_for(NvNv, c->nsm), 256, 0, c->stream>>>(dev, d, (int)(k % c->No), (int)(k / c->No),
                                                                         c->VIJ, c->vtab);
    else
      ingest_Vabij_kernel<<<grid_for(NvNv, c->nsm), 256, 0, c->stream>>>(dev, d, (int)(k % c->No), (int)(k / c->No),
                                                                       c->VIJ, c->vtab);
  });
  CUDA_OK(cudaGetLastError());
}

void load_ppph_impl(atrip_b200_ctx *c, const double *V, double *BY) {
  const StoreDims d = dims_of(c);
  const size_t NvNv = (size_t)c->Nv * c->Nv;
  // chunk = up to 16 consecutive E for one r: Vabci[:, :, E0.., r]; E blocks never straddle r
  const int Nv = c->Nv;
  const int eblocks = (Nv + KC - 1) / KC;
  const size_t z = esz(c);
  ensure_stage(c, z * (size_t)KC * NvNv);
  const dim3 grid((unsigned)((NvNv + 31) / 32)), block(32, 8);
  int slot = 0;
  for (int r = 0; r < c->No; r++)
    for (int eb = 0; eb < eblocks; eb++, slot ^= 1) {
      const int E0 = eb * KC, ne = std::min(KC, Nv - E0);
      const double *src = V + z * ((size_t)E0 + (size_t)r * Nv) * NvNv;
      const size_t n = z * (size_t)ne * NvNv;
      CUDA_OK(cudaEventSynchronize(c->stage_ev[slot]));
      if (!is_pinned(src)) {
        std::memcpy(c->h_stage[slot], src, n * 8);
        src = c->h_stage[slot];
      }
      CUDA_OK(cudaMemcpyAsync(c->d_stage[slot], src, n * 8, cudaMemcpyHostToDevice, c->stream));
      if (c->cplx)
        ingest_Vabci_z_kernel<<<grid_for((size_t)ne * NvNv, c->nsm), 256, 0, c->stream>>>(c->d_stage[slot], d, E0, ne, r,
                                                                                        BY, c->btab);
      else
        ingest_Vabci_kernel<<<grid, block, 0, c->stream>>>(c->d_stage[slot], d, E0, ne, r, BY, c->btab);
      CUDA_OK(cudaEventRecord(c->stage_ev[slot], c->stream));
    }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaGetLastError());
}


// ------------------------------------------------------------------ per-slice ingest / read-back
size_t ref_slice_elems(const atrip_b200_ctx *c, int kind) {  // elements (F) of a slice in the reference layout
  const size_t No = c->No, Nv = c->Nv;
  switch (kind) {
    case 100: return Nv * No * No;
    case 101: case 111: return No * No * No;
    case 200: case 210: return Nv * No;
    case 201: case 202: return No * No;
  }
  throw Fail{"unknown slice kind"};
}

void check_xy(const atrip_b200_ctx *c, int kind, int64_t n, const int64_t *xy) {
  const int64_t Nv = c->Nv;
  for (int64_t i = 0; i < n; i++) {
    const int64_t x = xy[2 * i], y = xy[2 * i + 1];
    REQUIRE(x >= 0 && x < Nv, "slice index x out of range");
    if (kind >= 200) REQUIRE(y >= 0 && y < Nv, "slice index y out of range");
    if (kind == 201 || kind == 202) REQUIRE(x <= y, "TABIJ / VABIJ slices are addressed with x <= y");
  }
}

// Replaces SliceUnion<F>::init (SliceUnion.cxx:305-332) for one kind: n slices in the reference's
// slice layout, back to back in host memory, go through the staging pair in chunks and are
// scattered into the stores; slices this rank does not hold are skipped.  Pinned host memory is read
// by DMA straight from the caller's buffer (asynchronously: atrip_b200_upload_flush).
void upload_slices_impl(atrip_b200_ctx *c, int kind, int64_t n, const int64_t *xy, const double *host) {
  const size_t per = ref_slice_elems(c, kind) * esz(c);  // doubles per slice
  REQUIRE(n >= 0, "negative slice count");
  if (n == 0) return;
  check_xy(c, kind, n, xy);
  const bool J = kind == 111 || kind == 210;
  REQUIRE(!J || c->cfg.with_J, "context was created without with_J");
  begin_store_update(c);
  const int base_kind = kind == 111 ? 101 : (kind == 210 ? 200 : kind);
  const size_t chunk_slices = std::max<size_t>(1, (size_t)(48u << 20) / (per * 8));
  ensure_stage(c, chunk_slices * per);
  Scratch<long long> dxy((size_t)2 * std::min<size_t>(chunk_slices, (size_t)n) * 2);
  const bool pinned = is_pinned(host);
  const SliceTables tb{c->xtab, c->btab, c->vtab};
  const StoreDims d = dims_of(c);
  size_t k = 0;
  for (int64_t s0 = 0; s0 < n; s0 += (int64_t)chunk_slices, k++) {
    const size_t ns = (size_t)std::min<int64_t>((int64_t)chunk_slices, n - s0);
    const int st = (int)(k & 1);
    CUDA_OK(cudaEventSynchronize(c->stage_ev[st]));
    const double *src = host + (size_t)s0 * per;
    if (!pinned) {
      std::memcpy(c->h_stage[st], src, ns * per * 8);
      src = c->h_stage[st];
    }
    long long *dx = dxy.p + (size_t)st * 2 * std::min<size_t>(chunk_slices, (size_t)n);
    CUDA_OK(cudaMemcpyAsync(dx, xy + 2 * s0, ns * 2 * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_stage[st], src, ns * per * 8, cudaMemcpyHostToDevice, c->stream));
    const int grid = grid_for(ns * ref_slice_elems(c, kind), c->nsm);
    auto launch = [&](double *AX, double *BY) {
      if (c->cplx) upload_slices_kernel<true><<<grid, 256, 0, c->stream>>>(base_kind, c->d_stage[st], dx, (int)ns, d, tb, AX, BY, c->VIJ);
      else upload_slices_kernel<false><<<grid, 256, 0, c->stream>>>(base_kind, c->d_stage[st], dx, (int)ns, d, tb, AX, BY, c->VIJ);
    };
    if (J) launch(c->AXJ, c->BYJ);
    else {
      launch(c->AX, c->BY);
      // the J stores share the amplitudes (Tabij) with the V stores
      if (c->cfg.with_J && (kind == 100 || kind == 201)) launch(c->AXJ, c->BYJ);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(c->stage_ev[st], c->stream));
  }
  if (J) c->have_J = true;
  // pageable sources were consumed by the memcpy above; the device work may still be in flight
  // (the slot tables dxy are freed below: wait for the kernels that read them)
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

void read_slices_impl(atrip_b200_ctx *c, int kind, int64_t n, const int64_t *xy, double *out) {
  const size_t per = ref_slice_elems(c, kind) * esz(c);
  REQUIRE(n >= 0, "negative slice count");
  if (n == 0) return;
  check_xy(c, kind, n, xy);
  const bool J = kind == 111 || kind == 210;
  REQUIRE(!J || c->cfg.with_J, "context was created without with_J");
  const int base_kind = kind == 111 ? 101 : (kind == 210 ? 200 : kind);
  const size_t chunk_slices = std::max<size_t>(1, (size_t)(48u << 20) / (per * 8));
  Scratch<double> dbuf(std::min<size_t>(chunk_slices, (size_t)n) * per);
  Scratch<long long> dxy((size_t)2 * std::min<size_t>(chunk_slices, (size_t)n));
  const SliceTables tb{c->xtab, c->btab, c->vtab};
  const StoreDims d = dims_of(c);
  for (int64_t s0 = 0; s0 < n; s0 += (int64_t)chunk_slices) {
    const size_t ns = (size_t)std::min<int64_t>((int64_t)chunk_slices, n - s0);
    CUDA_OK(cudaMemcpyAsync(dxy.p, xy + 2 * s0, ns * 2 * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    const int grid = grid_for(ns * ref_slice_elems(c, kind), c->nsm);
    const double *AX = J ? c->AXJ : c->AX, *BY = J ? c->BYJ : c->BY;
    if (c->cplx) read_slices_kernel<true><<<grid, 256, 0, c->stream>>>(base_kind, dbuf.p, dxy.p, (int)ns, d, tb, AX, BY, c->VIJ);
    else read_slices_kernel<false><<<grid, 256, 0, c->stream>>>(base_kind, dbuf.p, dxy.p, (int)ns, d, tb, AX, BY, c->VIJ);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(out + (size_t)s0 * per, dbuf.p, ns * per * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
}

// ------------------------------------------------------------------ slice exchange (NCCL)
#define NCCL_OK(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      throw Fail{std::string("NCCL: ") + nccl().GetErrorString(r__) + " in " #expr " (" __FILE__ ":" + \
                 std::to_string(__LINE__) + ")"};                                                  \
  } while (0)

int32_t *req_buf(const atrip_b200_ctx *c, int32_t *base, int parity, int peer) {
  return base + ((size_t)parity * c->cfg.nranks + peer) * c->req_cap;
}

double *store_of(const atrip_b200_ctx *c, int kind, bool J) {
  return kind == KA ? (J ? c->AXJ : c->AX) : (kind == KB ? (J ? c->BYJ : c->BY) : c->VIJ);
}
double *cache_of(const atrip_b200_ctx *c, int kind, bool J) {
  return kind == KA ? (J ? c->cAJ : c->cA) : (kind == KB ? (J ? c->cBJ : c->cB) : c->cV);
}

// One exchange step on the side stream (replaces do_io_phase, Atrip.cxx:461-572, one BATCH at a
// time): inside one NCCL group
//   data     for batch `k`:  send the ranges every peer asked of me (peer_req, host copy of the
//            request lists received one step earlier), receive the ranges of my own plan into
//            the cache slots the plan assigned;
//   requests for batch k+1:  send my request lists (next != nullptr), receive the peers'.
// k = -1 is the bootstrap step that only exchanges the request lists of batch 0.
void exchange_step(atrip_b200_ctx *c, int64_t k, const BatchPlan *mine, const BatchPlan *next) {
  NcclApi &N = nccl();
  const int n = c->cfg.nranks, me = c->cfg.rank;
  const bool J = c->have_J;
  const int nextpar = (int)((k + 1) & 1), par = (int)(k & 1);
  if (next) {
    for (int p = 0; p < n; p++) encode_requests(next->fetch[(size_t)p], req_buf(c, c->h_req_send, nextpar, p));
    CUDA_OK(cudaMemcpyAsync(req_buf(c, c->d_req_send, nextpar, 0), req_buf(c, c->h_req_send, nextpar, 0),
                            (size_t)n * c->req_cap * sizeof(int32_t), cudaMemcpyHostToDevice, c->xstream));
  }
  NCCL_OK(N.GroupStart());
  for (int p = 0; p < n; p++) {
    if (p == me) continue;
    if (k >= 0) {
      // what peer p asked of me for its batch k, in its order
      const int32_t *rq = req_buf(c, c->h_req_recv, par, p);
      const int nr = rq[0];
      REQUIRE(nr >= 0 && (size_t)(1 + 3 * (int64_t)nr) <= c->req_cap, "corrupt slice request list");
      for (int i = 0; i < nr; i++) {
        const int kind = rq[1 + 3 * i];
        const int64_t slot = rq[2 + 3 * i], cnt = rq[3 + 3 * i];
        REQUIRE(kind >= 0 && kind < 3 && slot >= 0 && cnt > 0 && slot + cnt <= c->owned[kind],
                "peer requested a slice this rank does not own");
        const size_t el = slice_elems(c, kind);
        NCCL_OK(N.Send(store_of(c, kind, false) + (size_t)slot * el, (size_t)cnt * el, ncclFloat64, p, c->comm, c->xstream));
        if (J && kind != KV)
          NCCL_OK(N.Send(store_of(c, kind, true) + (size_t)slot * el, (size_t)cnt * el, ncclFloat64, p, c->comm, c->xstream));
      }
      for (const FetchRange &fr : mine->fetch[(size_t)p]) {
        const size_t el = slice_elems(c, fr.kind);
        const size_t dst = (size_t)fr.dst_slot * el;
        NCCL_OK(N.Recv(cache_of(c, fr.kind, false) + dst, (size_t)fr.count * el, ncclFloat64, p, c->comm, c->xstream));
        if (J && fr.kind != KV)
          NCCL_OK(N.Recv(cache_of(c, fr.kind, true) + dst, (size_t)fr.count * el, ncclFloat64, p, c->comm, c->xstream));
        c->exch_bytes += (double)fr.count * el * 8 * ((J && fr.kind != KV) ? 2 : 1);
        c->exch_msgs += 1;
      }
    }
    if (next) {
      NCCL_OK(N.Send(req_buf(c, c->d_req_send, nextpar, p), c->req_cap, ncclInt32, p, c->comm, c->xstream));
      NCCL_OK(N.Recv(req_buf(c, c->d_req_recv, nextpar, p), c->req_cap, ncclInt32, p, c->comm, c->xstream));
    }
  }
  NCCL_OK(N.GroupEnd());
  if (next)
    CUDA_OK(cudaMemcpyAsync(req_buf(c, c->h_req_recv, nextpar, 0), req_buf(c, c->d_req_recv, nextpar, 0),
                            (size_t)n * c->req_cap * sizeof(int32_t), cudaMemcpyDeviceToHost, c->xstream));
}

// all ranks have finished filling their stores before anybody pulls from a peer
void rank_barrier(atrip_b200_ctx *c) {
  if (c->cfg.nranks == 1 || c->solo) return;
  REQUIRE(c->comm, "sharded stores need atrip_b200_comm_init before running tuples");
  CUDA_OK(cudaMemsetAsync(c->d_reduce, 0, sizeof(double), c->stream));
  NCCL_OK(nccl().AllReduce(c->d_reduce, c->d_reduce, 1, ncclFloat64, ncclSum, c->comm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

// P2P transport: pull the remote slices of one batch straight out of the owners' stores into the
// cache slots of the plan with the copy engines (NVLink / NVSwitch); nothing runs on an SM and the
// owner is not involved.  Replaces SliceUnion::receive + send (SliceUnion.cxx:365-505).  The AX / BY
// copies are dealt round-robin to NXS streams (a single stream serialises them: ~154 copies of 2.4 MB
// per c2 batch took the whole batch time at 8 ranks); c->xstream is ordered behind all of them again.
// Two orderings, because the slots a copy overwrites were last addressed two batches earlier
// (SliceCache): AX / BY slots are read by the contraction only, so their copies wait for
// `after_contract` (contraction of that batch done); Vabij slots are read by the reduction only and
// travel on their own stream c->xv behind `after_reduce`.  (With one ordering behind the reduction the
// copies of batch k+1 could only start when the low-priority reduction of batch k-1 had run, i.e. when
// the contraction of batch k drained -- exactly when the compute stream wants them.)
void pull_step(atrip_b200_ctx *c, const BatchPlan *mine, cudaEvent_t after_contract, cudaEvent_t after_reduce) {
  const bool J = c->have_J;
  if (after_contract) CUDA_OK(cudaStreamWaitEvent(c->xstream, after_contract, 0));
  if (after_reduce) CUDA_OK(cudaStreamWaitEvent(c->xv, after_reduce, 0));
  size_t ncopies = 0;
  for (const auto &v : mine->fetch)
    for (const FetchRange &fr : v) ncopies += fr.kind != KV;
  bool any = false;
  for (const auto &v : mine->fetch) any = any || !v.empty();
  if (!any) return;
  REQUIRE(!c->solo, "ATRIP_B200_SOLO_SHARD: a tuple reads a slice this rank does not own");
  const int ns = (int)std::max<size_t>(1, std::min<size_t>(NXS, ncopies));
  if (ns > 1) {
    CUDA_OK(cudaEventRecord(c->xfork, c->xstream));
    for (int i = 1; i < ns; i++) CUDA_OK(cudaStreamWaitEvent(c->xs[i], c->xfork, 0));
  }
  size_t i = 0;
  for (int p = 0; p < c->cfg.nranks; p++) {
    if (p == c->cfg.rank) continue;
    for (const FetchRange &fr : mine->fetch[(size_t)p]) {
      const size_t el = slice_elems(c, fr.kind);
      const size_t dst = (size_t)fr.dst_slot * el, src = (size_t)fr.src_slot * el;
      const size_t bytes = (size_t)fr.count * el * sizeof(double);
      cudaStream_t st = fr.kind == KV ? c->xv : c->xs[i++ % ns];
      CUDA_OK(cudaMemcpyAsync(cache_of(c, fr.kind, false) + dst, c->peer[fr.kind][(size_t)p] + src, bytes,
                              cudaMemcpyDeviceToDevice, st));
      if (J && fr.kind != KV)
        CUDA_OK(cudaMemcpyAsync(cache_of(c, fr.kind, true) + dst, c->peer[3 + fr.kind][(size_t)p] + src, bytes,
                                cudaMemcpyDeviceToDevice, st));
      c->exch_bytes += (double)bytes * ((J && fr.kind != KV) ? 2 : 1);
      c->exch_msgs += 1;
    }
  }
  for (int q = 1; q < ns; q++) {
    CUDA_OK(cudaEventRecord(c->xjoin[q], c->xs[q]));
    CUDA_OK(cudaStreamWaitEvent(c->xstream, c->xjoin[q], 0));
  }
}

// Runs the tuples list[0..count) in device batches.  Replaces the main loop, Atrip.cxx:686-1057.
// Sharded stores: COLLECTIVE with the NCCL transport -- every rank calls with the same count;
// slices of batch k+1 travel on the side stream(s) while batch k computes.  `next` (may be null):
// the tuples the caller is likely to run next (the list continues there); with the P2P transport
// their remote slices are prefetched behind the last batch, so that the next call does not start
// with an exposed fetch.
void run_list(atrip_b200_ctx *c, const Tuple *list, int64_t count, double *energy, double *ct_energy,
              const Tuple *next = nullptr, int64_t next_count = 0) {
  const bool ct = c->have_J;
  const bool sh = sharded(c);
  REQUIRE(!sh || c->comm || c->solo, "sharded stores need atrip_b200_comm_init before running tuples");
  REQUIRE(!sh || c->transport != 2 || !c->peer[0].empty() || c->solo, "peer stores are not mapped");
  const int64_t nb = (count + c->batch - 1) / c->batch;
  auto nt_of = [&](int64_t k) { return (size_t)std::min<int64_t>(c->batch, count - k * c->batch); };
  const bool p2p = sh && c->transport == 2;
  if (sh && c->stores_dirty) {
    rank_barrier(c);  // every rank has finished (re)filling its stores before anybody reads a peer's
    c->stores_dirty = false;
    c->cache.invalidate();
  }
  // everything of earlier runs has completed (each run ends with a full synchronisation): every
  // cache slot may be re-assigned, and what the slots hold stays addressable
  c->serial += 2;
  const int64_t serial0 = c->serial;
  BatchPlan plans[3];
  double plan_ms = 0, hits = 0, misses = 0;
  auto plan_tuples = [&](const Tuple *t, size_t n, int64_t serial, BatchPlan &pl, bool speculative) {
    const auto h0 = std::chrono::steady_clock::now();
    plan_batch(c->map, t, n, c->owned, c->cache, serial, pl);
    plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
    if (pl.overflow) {
      c->cache.invalidate();  // the plan is half applied: drop every key (nothing in flight addresses them by key)
      REQUIRE(speculative, "fetch cache too small for this batch (tuple list changed without set_tuples?)");
      return false;
    }
    for (int q = 0; q < 3; q++) {
      hits += (double)pl.hits[q];
      misses += (double)pl.used[q];
    }
    return true;
  };
  auto make_plan = [&](int64_t k) { plan_tuples(list + k * c->batch, nt_of(k), serial0 + k, plans[k % 3], false); };
  c->serial = serial0 + nb;
  c->exch_bytes = c->exch_msgs = 0;
  CUDA_OK(cudaMemsetAsync(c->d_total, 0, 2 * sizeof(double), c->stream));
  double ms_gap = 0, ms_contract = 0, ms_reduce = 0;
  int n_contract = 0, n_reduce = 0;
  int64_t harvested = 0;
  auto harvest = [&](int64_t k) {  // events of batch k have completed
    const BatchTimers &b = c->bt[k % TRING];
    float g = 0, a = 0, r = 0;
    CUDA_OK(cudaEventElapsedTime(&g, b.w0, b.c0));
    CUDA_OK(cudaEventElapsedTime(&a, b.c0, b.c1));
    CUDA_OK(cudaEventElapsedTime(&r, b.r0, b.r1));
    if (k > 0) ms_gap += g;  // batch 0: the gap is the start-up of the call, reported separately
    else c->timing[13] = g;
    ms_contract += a;
    ms_reduce += r;
  };
  CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
  if (nb > 0) make_plan(0);
  if (sh && nb > 0) {
    if (p2p) {
      pull_step(c, &plans[0], nullptr, nullptr);
      CUDA_OK(cudaEventRecord(c->xvdone[0], c->xv));
    } else {
      exchange_step(c, -1, nullptr, &plans[0]);
      CUDA_OK(cudaStreamSynchronize(c->xstream));
      if (nb > 1) make_plan(1);
      exchange_step(c, 0, &plans[0], nb > 1 ? &plans[1] : nullptr);
    }
    CUDA_OK(cudaEventRecord(c->xdone[0], c->xstream));
  }
  for (int64_t k = 0; k < nb; k++) {
    const int nt = (int)nt_of(k);
    const BatchPlan &pl = plans[k % 3];
    // ---- records of the batch -> device (ring of REC_RING pinned slots)
    const int slot = (int)(c->rec_uses % REC_RING);
    if (c->rec_uses >= REC_RING) CUDA_OK(cudaEventSynchronize(c->rec_ev[slot]));
    c->rec_uses++;
    while (harvested + TRING <= k) harvest(harvested++);  // batch k - TRING is long done (REC_RING < TRING)
    const BatchTimers &bt = c->bt[k % TRING];
    TupleRec *hr = c->h_recs + (size_t)slot * c->batch, *dr = c->d_recs + (size_t)slot * c->batch;
    std::memcpy(hr, pl.recs.data(), sizeof(TupleRec) * nt);
    CUDA_OK(cudaEventRecord(bt.w0, c->stream));
    if (sh) CUDA_OK(cudaStreamWaitEvent(c->stream, c->xdone[k & 3], 0));
    if (p2p) CUDA_OK(cudaStreamWaitEvent(c->rstream, c->xvdone[k & 3], 0));  // Vabij blocks of batch k
    CUDA_OK(cudaMemcpyAsync(dr, hr, sizeof(TupleRec) * nt, cudaMemcpyHostToDevice, c->stream));
    // ---- contraction of batch k on the high-priority stream into cube buffer k % 2 ...
    //      (the two kernels cannot share an SM profitably: plain FP64 instructions and DMMA use
    //      the same datapath, profiles/r01_fused_reducers_rejected_ncu.txt; the second stream only
    //      lets the reduction fill the ragged tail of the next contraction launch)
    const int buf = (int)(k & 1);
    c->last_nt = nt;
    c->last_buf = buf;
    if (k >= 2) CUDA_OK(cudaStreamWaitEvent(c->stream, c->rdone[(k - 2) & 3], 0));  // buffer reduced
    CUDA_OK(cudaEventRecord(bt.c0, c->stream));
    for (int var = 0; var <= c->cplx; var++) {  // complex field: Re cubes, then Im cubes
      launch_contract(c, dr, nt, false, buf, var);
      n_contract++;
    }
    CUDA_OK(cudaEventRecord(bt.c1, c->stream));
    CUDA_OK(cudaEventRecord(c->evV[k & 3], c->stream));
    if (ct) {
      for (int var = 0; var <= c->cplx; var++) {
        launch_contract(c, dr, nt, true, buf, var);
        n_contract++;
      }
      CUDA_OK(cudaEventRecord(c->evJ[k & 3], c->stream));
    }
    CUDA_OK(cudaEventRecord(c->cdone[k & 3], c->stream));
    // ---- ... and its reduction on the low-priority stream
    CUDA_OK(cudaStreamWaitEvent(c->rstream, c->evV[k & 3], 0));
    CUDA_OK(cudaEventRecord(bt.r0, c->rstream));
    launch_reduce(c, dr, nt, false, buf, c->d_total);
    n_reduce += 2;
    CUDA_OK(cudaEventRecord(bt.r1, c->rstream));
    if (ct) {
      CUDA_OK(cudaStreamWaitEvent(c->rstream, c->evJ[k & 3], 0));
      launch_reduce(c, dr, nt, true, buf, c->d_total + 1);
      n_reduce += 2;
    }
    CUDA_OK(cudaEventRecord(c->rdone[k & 3], c->rstream));
    CUDA_OK(cudaEventRecord(c->rec_ev[slot], c->rstream));
    // ---- next batch: host plan (and, sharded, its exchange on the side stream).  The copies of
    //      batch k+1 may overwrite slots last used by batch k-1: they wait for its reduction.
    if (k + 1 < nb) {
      if (!sh) make_plan(k + 1);
      else if (p2p) {  // fully asynchronous: the host never waits for a transfer
        make_plan(k + 1);
        pull_step(c, &plans[(k + 1) % 3], k >= 1 ? c->cdone[(k - 1) & 3] : nullptr, k >= 1 ? c->rdone[(k - 1) & 3] : nullptr);
        CUDA_OK(cudaEventRecord(c->xdone[(k + 1) & 3], c->xstream));
        CUDA_OK(cudaEventRecord(c->xvdone[(k + 1) & 3], c->xv));
      } else {
        CUDA_OK(cudaEventSynchronize(c->xdone[k & 3]));  // peers' requests for batch k+1 are on the host
        if (k + 2 < nb) make_plan(k + 2);
        if (k >= 1) CUDA_OK(cudaStreamWaitEvent(c->xstream, c->rdone[(k - 1) & 3], 0));
        exchange_step(c, k + 1, &plans[(k + 1) % 3], k + 2 < nb ? &plans[(k + 2) % 3] : nullptr);
        CUDA_OK(cudaEventRecord(c->xdone[(k + 1) & 3], c->xstream));
      }
    } else if (p2p && next && next_count > 0) {
      // prefetch for the next call: the slices of the batch that follows this list in the caller's
      // tuple list go to the cache now, behind the last batch's compute
      BatchPlan &pl2 = plans[(k + 1) % 3];
      if (plan_tuples(next, (size_t)std::min<int64_t>(c->batch, next_count), serial0 + nb, pl2, true)) {
        c->serial = serial0 + nb + 1;
        pull_step(c, &pl2, k >= 1 ? c->cdone[(k - 1) & 3] : nullptr, k >= 1 ? c->rdone[(k - 1) & 3] : nullptr);
      }
    }
  }
  if (nb > 0) CUDA_OK(cudaStreamWaitEvent(c->stream, c->rdone[(nb - 1) & 3], 0));  // the last reduction
  CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
  double tot[2];
  CUDA_OK(cudaMemcpyAsync(tot, c->d_total, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaStreamSynchronize(c->rstream));
  if (sh) {
    for (int i = 0; i < NXS; i++) CUDA_OK(cudaStreamSynchronize(c->xs[i]));
    CUDA_OK(cudaStreamSynchronize(c->xv));
  }
  while (harvested < nb) harvest(harvested++);
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  int64_t real = 0;
  for (int64_t t = 0; t < count; t++) real += !is_fake(list[t]);
  c->timing[0] = ms;
  c->timing[1] = nb ? ms_contract / nb : 0;  // mean ms per batch: contraction launch(es) of the V pass
  c->timing[2] = nb ? ms_reduce / nb : 0;    // mean ms per batch: reduction (+ batch sum) of the V pass
  c->timing[3] = n_contract;
  c->timing[4] = n_reduce;
  c->timing[5] = (double)real;
  c->timing[6] = c->exch_bytes;
  c->timing[7] = c->exch_msgs;
  c->timing[8] = ms_gap;     // device: sum over batches 1.. of the idle gap in front of the contraction launch
  c->timing[9] = plan_ms;    // host: building the slot records / fetch schedule
  c->timing[10] = hits;      // remote slices found in the cache
  c->timing[11] = misses;    // remote slices fetched
  c->timing[12] = (double)nb;
  if (energy) *energy = tot[0];
  if (ct_energy) *ct_energy = ct ? tot[1] : tot[0];  // without J the reference's ct_energy == energy
}

void run_impl(atrip_b200_ctx *c, int64_t first, int64_t count, double *energy, double *ct_energy) {
  REQUIRE(first >= 0 && count >= 0 && (size_t)(first + count) <= c->tuples.size(), "tuple range out of bounds");
  const int64_t rest = (int64_t)c->tuples.size() - (first + count);
  run_list(c, c->tuples.data() + first, count, energy, ct_energy, rest > 0 ? c->tuples.data() + first + count : nullptr, rest);
}

void set_tuples_impl(atrip_b200_ctx *c) {
  int64_t need[3] = {0, 0, 0};
  if (sharded(c)) cache_need(c->map, c->tuples.data(), c->tuples.size(), (size_t)c->batch, need);
  ensure_caches(c, need);
}

void tuple_debug_impl(atrip_b200_ctx *c, int64_t a, int64_t b, int64_t cc, double *Tijk, double *Zijk, double *energy) {
  REQUIRE(a >= 0 && a <= b && b <= cc && cc < c->Nv && !(a == b && b == cc), "not a valid tuple a<=b<=c");
  const Tuple one{(uint64_t)a, (uint64_t)b, (uint64_t)cc};
  double e = 0;
  const uint64_t slot = c->rec_uses % REC_RING;  // the ring slot run_list is about to use
  run_list(c, &one, 1, &e, nullptr);
  const size_t cube = (size_t)c->No * c->No * c->No, cd = esz(c) * cube;  // complex: interleaved cubes
  Scratch<double> sT, sZ;
  if (Tijk) sT.alloc(cd);
  if (Zijk) sZ.alloc(cd);
  double *dT = sT.p, *dZ = sZ.p;
  if (Tijk || Zijk) {
    ReduceParams P = reduce_params(c, c->d_recs + slot * c->batch, 1, false, 0);
    if (c->cplx) cubes_z_kernel<<<grid_for(cube, c->nsm), 256, 0, c->stream>>>(P, 0, dT, dZ);
    else cubes_kernel<<<grid_for(cube, c->nsm), 256, 0, c->stream>>>(P, 0, dT, dZ);
    CUDA_OK(cudaGetLastError());
  }
  if (Tijk) CUDA_OK(cudaMemcpyAsync(Tijk, dT, cd * 8, cudaMemcpyDeviceToHost, c->stream));
  if (Zijk) CUDA_OK(cudaMemcpyAsync(Zijk, dZ, cd * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (energy) *energy = e;
}

void read_slice_impl(atrip_b200_ctx *c, int kind, int64_t x, int64_t y, double *out) {
  const size_t No = c->No, Nv = c->Nv;
  const ShardMap &m = c->map;
  REQUIRE(x >= 0 && x < (int64_t)Nv, "slice index x out of range");
  size_t n = 0;
  const double *ax = nullptr, *by = nullptr, *vij = nullptr;
  if (kind == 100 || kind == 101 || kind == 201) {
    n = kind == 100 ? Nv * No * No : (kind == 101 ? No * No * No : No * No);
    REQUIRE(m.ownerA(x) == m.me, "this rank does not own that slice");
    ax = c->AX + (size_t)m.slotA(x) * slice_elems(c, KA);
  } else if (kind == 200) {
    REQUIRE(y >= 0 && y < (int64_t)Nv, "slice index y out of range");
    n = Nv * No;
    const int64_t id = m.idB(x, y, false);
    REQUIRE(m.ownerB(id) == m.me, "this rank does not own that slice");
    by = c->BY + (size_t)m.slotB(id) * slice_elems(c, KB);
  } else if (kind == 202) {
    REQUIRE(y >= x && y < (int64_t)Nv, "VABIJ slices are stored for x <= y");
    n = No * No;
    const int64_t s = m.localV(x, y);
    REQUIRE(s >= 0, "this rank does not own that slice");
    vij = c->VIJ + (size_t)s * slice_elems(c, KV);
  } else {
    throw Fail{"unknown slice kind"};
  }
  if (kind == 201) REQUIRE(y >= 0 && y < (int64_t)Nv, "slice index y out of range");
  Scratch<double> sd(esz(c) * n);
  double *d = sd.p;
  if (c->cplx) read_slice_z_kernel<<<grid_for(n, c->nsm), 256, 0, c->stream>>>(kind, dims_of(c), ax, by, vij, (int)y, d);
  else read_slice_kernel<<<grid_for(n, c->nsm), 256, 0, c->stream>>>(kind, dims_of(c), ax, by, vij, (int)y, d);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(out, d, esz(c) * n * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

void comm_init_impl(atrip_b200_ctx *c, const void *id128) {
  NcclApi &N = nccl();
  REQUIRE(N.ok, "NCCL is not available: " + N.error);
  REQUIRE(!c->comm, "communicator already initialised");
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(&id, id128, sizeof(id));
  NCCL_OK(N.CommInitRank(&c->comm, c->cfg.nranks, id, c->cfg.rank));
  if (!sharded(c) || c->transport != 2) return;
  // P2P transport: every rank publishes CUDA IPC handles of its stores (all-gather over the new
  // communicator) and maps the peers' stores into its own address space
  const int n = c->cfg.nranks, me = c->cfg.rank;
  double *mine[5] = {c->AX, c->BY, c->VIJ, c->AXJ, c->BYJ};
  const int nh = c->cfg.with_J ? 5 : 3;
  std::vector<cudaIpcMemHandle_t> all((size_t)n * 5);
  for (int k = 0; k < nh; k++) CUDA_OK(cudaIpcGetMemHandle(&all[(size_t)me * 5 + k], mine[k]));
  const size_t per = 5 * sizeof(cudaIpcMemHandle_t);
  Scratch<unsigned char> sd(per * n);
  unsigned char *d = sd.p;
  CUDA_OK(cudaMemcpyAsync(d + per * me, &all[(size_t)me * 5], per, cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(N.GroupStart());
  for (int p = 0; p < n; p++) {
    if (p == me) continue;
    NCCL_OK(N.Send(d + per * me, per, ncclUint8, p, c->comm, c->stream));
    NCCL_OK(N.Recv(d + per * p, per, ncclUint8, p, c->comm, c->stream));
  }
  NCCL_OK(N.GroupEnd());
  CUDA_OK(cudaMemcpyAsync(all.data(), d, per * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 5; k++) c->peer[k].assign((size_t)n, nullptr);
  for (int p = 0; p < n; p++)
    for (int k = 0; k < nh; k++) {
      if (p == me) { c->peer[k][(size_t)p] = mine[k]; continue; }
      void *ptr = nullptr;
      CUDA_OK(cudaIpcOpenMemHandle(&ptr, all[(size_t)p * 5 + k], cudaIpcMemLazyEnablePeerAccess));
      c->peer[k][(size_t)p] = (double *)ptr;
    }
}

void allreduce_impl(atrip_b200_ctx *c, double *vals, int n) {
  REQUIRE(n >= 0 && n <= 16, "at most 16 values");
  if (c->cfg.nranks == 1 || n == 0) return;
  REQUIRE(c->comm, "atrip_b200_allreduce needs atrip_b200_comm_init");
  CUDA_OK(cudaMemcpyAsync(c->d_reduce, vals, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(nccl().AllReduce(c->d_reduce, c->d_reduce, (size_t)n, ncclFloat64, ncclSum, c->comm, c->stream));
  CUDA_OK(cudaMemcpyAsync(vals, c->d_reduce, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
}

// register-resident DMMA loop: the FP64 tensor ceiling (same loop as tools/fp64_peak.cu)
__global__ void dmma_peak_kernel(double *out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 0.0;
  const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void synth_range_kernel(double *out, uint64_t key, int tensor_id, double scale, uint64_t first,
                                   uint64_t count) {
  for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < count; e += (uint64_t)gridDim.x * blockDim.x) {
    const double u = synth_u(key, first + e);
    out[e] = tensor_id == T_EPS_I   ? __dadd_rn(-2.0, __dmul_rn(1.5, u))
             : tensor_id == T_EPS_A ? __dadd_rn(0.5, __dmul_rn(3.5, u))
                                    : __dmul_rn(scale, __dadd_rn(u, -0.5));
  }
}

// debug: order-independent checksum (sum of the raw bit patterns as integers) of n doubles
__global__ void checksum_kernel(const double *p, size_t n, unsigned long long *out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += (unsigned long long)__double_as_longlong(p[i]);
  atomicAdd(out, s);
}

template <typename F>
int guarded(atrip_b200_ctx *c, F f) {
  try {
    if (c) CUDA_OK(cudaSetDevice(c->cfg.device));
    f();
    return 0;
  } catch (const Fail &e) {
    g_error = e.msg;
  } catch (const std::exception &e) {
    g_error = e.what();
  }
  return 1;
}

}  // namespace

extern "C" {

const char *atrip_b200_last_error(void) { return g_error.c_str(); }
const char *atrip_b200_version(void) { return "atrip_b200 0.2 (sm_100a)"; }
int32_t atrip_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int atrip_b200_create(atrip_b200_ctx **out, const atrip_b200_config *cfg) {
  if (!out || !cfg) {
    g_error = "null argument";
    return 1;
  }
  *out = nullptr;
  atrip_b200_ctx *c = new atrip_b200_ctx();
  c->cfg = *cfg;
  const int rc = guarded(nullptr, [&] { create_impl(c); });
  if (rc) {
    destroy_impl(c);
    return rc;
  }
  *out = c;
  return 0;
}

int atrip_b200_destroy(atrip_b200_ctx *c) {
  destroy_impl(c);
  return 0;
}

int atrip_b200_set_epsilon(atrip_b200_ctx *c, const double *ei, const double *ea) {
  return guarded(c, [&] {
    CUDA_OK(cudaMemcpy(c->eps_i, ei, sizeof(double) * esz(c) * c->No, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->eps_a, ea, sizeof(double) * esz(c) * c->Nv, cudaMemcpyHostToDevice));
  });
}
int atrip_b200_set_Tai(atrip_b200_ctx *c, const double *Tai) {
  return guarded(c, [&] { CUDA_OK(cudaMemcpy(c->Tai, Tai, sizeof(double) * esz(c) * c->No * c->Nv, cudaMemcpyHostToDevice)); });
}
int atrip_b200_load_Tabij(atrip_b200_ctx *c, const double *T) { return guarded(c, [&] { begin_store_update(c); load_Tabij_impl(c, T); }); }
int atrip_b200_load_Vabij(atrip_b200_ctx *c, const double *V) { return guarded(c, [&] { begin_store_update(c); load_Vabij_impl(c, V); }); }
int atrip_b200_load_Vijka(atrip_b200_ctx *c, const double *V) {
  return guarded(c, [&] { begin_store_update(c); load_hhhp_impl(c, V, c->AX); });
}
int atrip_b200_load_Vabci(atrip_b200_ctx *c, const double *V) {
  return guarded(c, [&] { begin_store_update(c); load_ppph_impl(c, V, c->BY); });
}
int atrip_b200_load_Jijka(atrip_b200_ctx *c, const double *V) {
  return guarded(c, [&] {
    REQUIRE(c->cfg.with_J, "context was created without with_J");
    begin_store_update(c);
    load_hhhp_impl(c, V, c->AXJ);
    c->have_J = true;
  });
}
int atrip_b200_load_Jabci(atrip_b200_ctx *c, const double *V) {
  return guarded(c, [&] {
    REQUIRE(c->cfg.with_J, "context was created without with_J");
    begin_store_update(c);
    load_ppph_impl(c, V, c->BYJ);
    c->have_J = true;
  });
}
int atrip_b200_fill_synthetic(atrip_b200_ctx *c, uint64_t seed, double scale) {
  return guarded(c, [&] { fill_impl(c, seed, scale); });
}

int atrip_b200_build_tuples(atrip_b200_ctx *c, int32_t distribution) {
  return guarded(c, [&] {
    REQUIRE(distribution == 0 || distribution == 1, "distribution must be 0 (NAIVE) or 1 (GROUP_AND_SORT)");
    c->tuples = distribution == 0 ? naive_tuples(c->Nv, c->cfg.rank, c->cfg.nranks)
                                  : group_and_sort_tuples(c->Nv, c->cfg.rank, c->cfg.nranks, true);
    set_tuples_impl(c);
  });
}
int atrip_b200_set_tuples(atrip_b200_ctx *c, const uint64_t *abc, int64_t n) {
  return guarded(c, [&] {
    c->tuples.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      const uint64_t a = abc[3 * i], b = abc[3 * i + 1], cc = abc[3 * i + 2];
      const bool fake = a == 0 && b == 0 && cc == 0;
      REQUIRE(fake || (a <= b && b <= cc && cc < (uint64_t)c->Nv && !(a == b && b == cc)),
              "tuple " + std::to_string(i) + " is not a<=b<=c<Nv (or the fake tuple)");
      c->tuples[i] = Tuple{a, b, cc};
    }
    set_tuples_impl(c);
  });
}
int64_t atrip_b200_num_tuples(const atrip_b200_ctx *c) { return c ? (int64_t)c->tuples.size() : 0; }
int atrip_b200_get_tuples(const atrip_b200_ctx *c, uint64_t *abc, int64_t cap) {
  const int64_t n = std::min<int64_t>(cap, (int64_t)c->tuples.size());
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) abc[3 * i + d] = c->tuples[i][d];
  return 0;
}

int atrip_b200_run(atrip_b200_ctx *c, int64_t first, int64_t count, double *energy, double *ct_energy) {
  return guarded(c, [&] { run_impl(c, first, count, energy, ct_energy); });
}
int atrip_b200_tuple_debug(atrip_b200_ctx *c, int64_t a, int64_t b, int64_t cc, double *T, double *Z, double *e) {
  return guarded(c, [&] { tuple_debug_impl(c, a, b, cc, T, Z, e); });
}
int atrip_b200_read_slice(atrip_b200_ctx *c, int32_t kind, int64_t x, int64_t y, double *out) {
  return guarded(c, [&] { read_slice_impl(c, kind, x, y, out); });
}
int atrip_b200_upload_slices(atrip_b200_ctx *c, int32_t kind, int64_t n, const int64_t *xy, const double *host) {
  return guarded(c, [&] { upload_slices_impl(c, kind, n, xy, host); });
}
int atrip_b200_upload_slice(atrip_b200_ctx *c, int32_t kind, int64_t x, int64_t y, const double *host) {
  const int64_t xy[2] = {x, y};
  return guarded(c, [&] { upload_slices_impl(c, kind, 1, xy, host); });
}
int atrip_b200_read_slices(atrip_b200_ctx *c, int32_t kind, int64_t n, const int64_t *xy, double *out) {
  return guarded(c, [&] { read_slices_impl(c, kind, n, xy, out); });
}
// slices of `kind` a rank has to be given (the sources it owns, SliceUnion.cxx:305-332 with
// RankMap::find, RankMap.cxx:35-85): (x, y) pairs in upload order; returns the count
int64_t atrip_b200_host_owned_slices(int32_t kind, int64_t Nv, int32_t rank, int32_t nranks, int64_t *xy, int64_t cap) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks) { g_error = "atrip_b200_host_owned_slices: bad arguments"; return -1; }
  int64_t k = 0;
  auto put = [&](int64_t x, int64_t y) {
    if (xy && k < cap) { xy[2 * k] = x; xy[2 * k + 1] = y; }
    k++;
  };
  if (kind == 100 || kind == 101 || kind == 111) {
    for (int64_t x = rank; x < Nv; x += nranks) put(x, 0);
  } else if (kind == 200 || kind == 210) {  // ordered pairs live with their first index
    for (int64_t x = rank; x < Nv; x += nranks)
      for (int64_t y = 0; y < Nv; y++) put(x, y);
  } else if (kind == 201 || kind == 202) {  // x <= y: held by the owner of x and by the owner of y
    for (int64_t y = 0; y < Nv; y++)
      for (int64_t x = 0; x <= y; x++)
        if (x % nranks == rank || y % nranks == rank) put(x, y);
  } else {
    g_error = "atrip_b200_host_owned_slices: unknown slice kind";
    return -1;
  }
  return k;
}
int atrip_b200_debug_cubes_checksum(atrip_b200_ctx *c, uint64_t *out) {
  return guarded(c, [&] {
    Scratch<unsigned long long> sd(1);
    unsigned long long *d = sd.p;
    CUDA_OK(cudaMemsetAsync(d, 0, 8, c->stream));
    const size_t n = (size_t)c->last_nt * ncubes(c) * cube_blocked_elems(c->No);
    if (n) checksum_kernel<<<1024, 256, 0, c->stream>>>(c->R[c->last_buf], n, d);
    CUDA_OK(cudaGetLastError());
    unsigned long long h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *out = h;
  });
}
int atrip_b200_last_timing(const atrip_b200_ctx *c, double *out6) {
  for (int i = 0; i < 6; i++) out6[i] = c->timing[i];
  return 0;
}
int atrip_b200_last_phases(const atrip_b200_ctx *c, double *out6) {
  for (int i = 0; i < 6; i++) out6[i] = c->timing[8 + i];
  return 0;
}
int atrip_b200_last_exchange(const atrip_b200_ctx *c, double *out2) {
  out2[0] = c->timing[6];
  out2[1] = c->timing[7];
  return 0;
}

int atrip_b200_comm_unique_id(void *id128) {
  return guarded(nullptr, [&] {
    NcclApi &N = nccl();
    REQUIRE(N.ok, "NCCL is not available: " + N.error);
    ncclUniqueId id;
    NCCL_OK(N.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
  });
}
int atrip_b200_comm_init(atrip_b200_ctx *c, const void *id128) {
  return guarded(c, [&] { comm_init_impl(c, id128); });
}
int atrip_b200_allreduce(atrip_b200_ctx *c, double *vals, int32_t n) {
  return guarded(c, [&] { allreduce_impl(c, vals, n); });
}
int64_t atrip_b200_kp(const atrip_b200_ctx *c) { return c->Kp; }
int64_t atrip_b200_batch_tuples(const atrip_b200_ctx *c) { return c->batch; }
double atrip_b200_flops_per_tuple(const atrip_b200_ctx *c) {
  const double No = c->No, Nv = c->Nv;
  return 12.0 * No * No * No * (No + Nv) * (c->cplx ? 4.0 : 1.0);  // x4 for complex, Atrip.cxx:578-580
}

int atrip_b200_measure_dmma_peak(int32_t device, double *tflops) {
  return guarded(nullptr, [&] {
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 2, threads = 256, iters = 40000;
    Scratch<double> sout((size_t)grid * threads);
    double *out = sout.p;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
      CUDA_OK(cudaEventRecord(e0));
      dmma_peak_kernel<<<grid, threads>>>(out, iters);
      CUDA_OK(cudaEventRecord(e1));
      CUDA_OK(cudaEventSynchronize(e1));
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = (double)grid * (threads / 32) * iters * 16 * 512.0 / (ms * 1e-3) / 1e12;
      if (rep > 0) best = std::max(best, tf);
    }
    CUDA_OK(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
  });
}

int atrip_b200_synth_to_host(int32_t device, uint64_t seed, int32_t tensor_id, double scale, uint64_t first,
                             uint64_t count, double *host) {
  return guarded(nullptr, [&] {
    CUDA_OK(cudaSetDevice(device));
    const uint64_t chunk = 1ull << 26;  // 512 MiB
    Scratch<double> sd(std::min<uint64_t>(chunk, std::max<uint64_t>(count, 1)));
    double *d = sd.p;
    for (uint64_t off = 0; off < count; off += chunk) {
      const uint64_t n = std::min(chunk, count - off);
      synth_range_kernel<<<2048, 256>>>(d, synth_key(seed, tensor_id), tensor_id, scale, first + off, n);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaMemcpy(host + off, d, n * sizeof(double), cudaMemcpyDeviceToHost));
    }
  });
}

int atrip_b200_host_plan(int64_t No, int64_t smem_limit_bytes, int64_t *out) {
  if (No < 1 || No > 256) {
    g_error = "atrip_b200_host_plan: No must be in [1, 256]";
    return 1;
  }
  const ContractPlan p = plan_contraction((int)No, smem_limit_bytes > 0 ? (size_t)smem_limit_bytes : 232448);
  if (!p.k) {
    g_error = "no contraction kernel variant fits";
    return 1;
  }
  const int64_t v[11] = {p.k->MI, p.k->NI, p.nw, p.tu, p.tv, p.mtiles, p.ntiles, p.nstages, (int64_t)p.smem,
                         p.arows, (int64_t)(p.useful * 1e6)};
  for (int i = 0; i < 11; i++) out[i] = v[i];
  return 0;
}

int64_t atrip_b200_host_tuples(int32_t distribution, int64_t Nv, int32_t rank, int32_t nranks, int32_t pad,
                               uint64_t *abc, int64_t cap) {
  if (Nv < 1 || nranks < 1 || rank < 0 || rank >= nranks || (distribution != 0 && distribution != 1)) {
    g_error = "atrip_b200_host_tuples: bad arguments";
    return -1;
  }
  if (distribution == 1 && (!abc || cap <= 0)) {  // length only: the container census suffices
    const gs::Census C = gs::census((uint64_t)Nv, (uint64_t)nranks);
    return (int64_t)(pad ? *std::max_element(C.cnt.begin(), C.cnt.end()) : C.cnt[(size_t)rank]);
  }
  std::vector<Tuple> t = distribution == 0 ? naive_tuples(Nv, rank, nranks)
                                           : group_and_sort_tuples(Nv, rank, nranks, pad != 0);
  if (distribution == 0 && !pad)
    while (!t.empty() && t.back() == Tuple{0, 0, 0}) t.pop_back();
  const int64_t n = (int64_t)t.size();
  for (int64_t i = 0; i < n && i < cap; i++)
    for (int d = 0; d < 3; d++) abc[3 * i + d] = t[i][d];
  return n;
}

namespace {
// reference slice kinds (Slice.hpp:99-108) -> store kind and id in that store
bool kind_to_store(int32_t kind, int64_t x, int64_t y, int64_t Nv, const ShardMap &m, int *store, int64_t *id) {
  if (x < 0 || x >= Nv) return false;
  if (kind == 100 || kind == 101) { *store = KA; *id = x; return true; }
  if (y < 0 || y >= Nv) return false;
  if (kind == 200 || kind == 201) { *store = KB; *id = m.idB(x, y, false); return true; }
  if (kind == 202 && x <= y) { *store = KV; *id = x + y * Nv; return true; }
  return false;
}
}  // namespace

int32_t atrip_b200_host_slice_slot(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t nranks, int64_t *slot) {
  if (nranks < 1 || Nv < 1) return -1;
  const ShardMap m(Nv, nranks, 0);
  int store;
  int64_t id;
  if (!kind_to_store(kind, x, y, Nv, m, &store, &id)) return -1;
  if (store == KA) { if (slot) *slot = m.slotA(id); return m.ownerA(id); }
  if (store == KB) { if (slot) *slot = m.slotB(id); return m.ownerB(id); }
  if (slot) *slot = m.slotV1(x, y);
  return m.ownerV(id);
}
int64_t atrip_b200_host_local_slot(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t rank, int32_t nranks) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks) return -1;
  const ShardMap m(Nv, nranks, rank);
  if (kind == 203) {
    if (x < 0 || x >= Nv) return -1;
    const int64_t id = m.idB(x, x, true);
    return m.ownerB(id) == rank ? m.slotB(id) : -1;
  }
  int store;
  int64_t id;
  if (!kind_to_store(kind, x, y, Nv, m, &store, &id)) return -1;
  if (store == KA) return m.ownerA(id) == rank ? m.slotA(id) : -1;
  if (store == KB) return m.ownerB(id) == rank ? m.slotB(id) : -1;
  return m.localV(x, y);
}
int32_t atrip_b200_host_slice_owner(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t nranks) {
  return atrip_b200_host_slice_slot(kind, x, y, Nv, nranks, nullptr);
}
int atrip_b200_host_shard_sizes(int64_t Nv, int32_t rank, int32_t nranks, int64_t *out3) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks) { g_error = "atrip_b200_host_shard_sizes: bad arguments"; return 1; }
  const ShardMap m(Nv, nranks, rank);
  for (int k = 0; k < 3; k++) out3[k] = m.owned(k, rank);
  return 0;
}
int64_t atrip_b200_host_plan_batch(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                                   const int64_t *cache_base3, int32_t *recs, int64_t *ranges, int64_t cap) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks || n < 0) { g_error = "atrip_b200_host_plan_batch: bad arguments"; return -1; }
  const ShardMap m(Nv, nranks, rank);
  std::vector<Tuple> t((size_t)n);
  for (int64_t i = 0; i < n; i++) t[i] = Tuple{abc[3 * i], abc[3 * i + 1], abc[3 * i + 2]};
  BatchPlan pl;
  SliceCache cache;  // empty cache large enough for any batch of n tuples: slots are taken in order 0, 1, ...
  const int64_t caps[3] = {3 * n + 1, 6 * n + 1, 3 * n + 1};
  cache.reset(caps);
  plan_batch(m, t.data(), (size_t)n, cache_base3, cache, 0, pl);
  if (recs) std::memcpy(recs, pl.recs.data(), sizeof(TupleRec) * (size_t)n);
  int64_t k = 0;
  for (int p = 0; p < nranks; p++)
    for (const FetchRange &fr : pl.fetch[(size_t)p]) {
      if (k < cap && ranges) {
        const int64_t v[5] = {p, fr.kind, fr.src_slot, fr.count, fr.dst_slot};
        std::memcpy(ranges + 5 * k, v, sizeof(v));
      }
      k++;
    }
  return k;
}

// Walks a whole tuple list in batches through the persistent cache, exactly as run_list does, and
// checks the invariants the engine relies on with an independent model of the cache contents:
//   * every cache slot a record addresses holds the slice the tuple needs at that point;
//   * the copies of batch k never overwrite a slot addressed by batch k or k - 1.
// calls: number of run calls the list is cut into (each call continues where the last one ended and
// prefetches the batch behind it, as the P2P transport does).  out[0..5] = slices fetched per kind,
// hits per kind; out[6] = copy ranges; out[7] = batches.  Returns 0, or -1 with last_error set.
int atrip_b200_host_check_schedule(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                                   int64_t batch, int32_t calls, const int64_t *cap3, double *out8) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks || n < 0 || batch < 1 || calls < 1) {
    g_error = "atrip_b200_host_check_schedule: bad arguments";
    return -1;
  }
  const ShardMap m(Nv, nranks, rank);
  std::vector<Tuple> t((size_t)n);
  for (int64_t i = 0; i < n; i++) t[i] = Tuple{abc[3 * i], abc[3 * i + 1], abc[3 * i + 2]};
  int64_t owned[3];
  for (int k = 0; k < 3; k++) owned[k] = m.owned(k, rank);
  SliceCache cache;
  cache.reset(cap3);
  // model: what each cache slot holds (peer, owner slot), and the last batch that addressed it
  std::vector<uint64_t> holds[3];
  std::vector<int64_t> used_by[3];
  for (int k = 0; k < 3; k++) {
    holds[k].assign((size_t)cap3[k], ~0ull);
    used_by[k].assign((size_t)cap3[k], INT64_MIN / 2);
  }
  double stats[8] = {0};
  int64_t serial = 0;
  BatchPlan pl;
  auto fail = [&](const std::string &msg) {
    g_error = "schedule check: " + msg;
    return -1;
  };
  auto apply = [&](const BatchPlan &p, int64_t ser) -> int {  // the copies of a batch
    for (int peer = 0; peer < nranks; peer++)
      for (const FetchRange &fr : p.fetch[(size_t)peer]) {
        stats[6] += 1;
        for (int64_t q = 0; q < fr.count; q++) {
          const int64_t s = fr.dst_slot + q;
          if (s < 0 || s >= cap3[fr.kind]) return fail("copy outside the cache");
          if (used_by[fr.kind][(size_t)s] >= ser - 1) return fail("copy overwrites a slot of a batch that may still compute");
          holds[fr.kind][(size_t)s] = ((uint64_t)peer << 48) | (uint64_t)(fr.src_slot + q);
        }
      }
    return 0;
  };
  auto verify = [&](const BatchPlan &p, const Tuple *tp, size_t nt, int64_t ser) -> int {
    for (size_t i = 0; i < nt; i++) {
      const TupleRec &r = p.recs[i];
      if (r.fake) continue;
      const int64_t x[3] = {(int64_t)tp[i][0], (int64_t)tp[i][1], (int64_t)tp[i][2]};
      auto check = [&](int kind, int slot, int owner, int64_t oslot) -> int {
        if (owner == rank) return slot == oslot ? 0 : 1;
        const int64_t cs = slot - owned[kind];
        if (cs < 0 || cs >= cap3[kind]) return 1;
        if (holds[kind][(size_t)cs] != (((uint64_t)owner << 48) | (uint64_t)oslot)) return 1;
        used_by[kind][(size_t)cs] = ser;
        return 0;
      };
      int bad = 0;
      for (int k = 0; k < 3; k++) bad += check(KA, r.ax[k], m.ownerA(x[k]), m.slotA(x[k]));
      const int64_t yz[6][3] = {{x[1], x[2], 0}, {x[0], x[2], 0}, {x[2], x[1], 1}, {x[0], x[1], 0}, {x[2], x[0], 1}, {x[1], x[0], 1}};
      for (int k = 0; k < 6; k++) {
        const int64_t id = m.idB(yz[k][0], yz[k][1], yz[k][2] != 0);
        bad += check(KB, r.by[k], m.ownerB(id), m.slotB(id));
      }
      const int64_t vp[3][2] = {{x[1], x[2]}, {x[0], x[2]}, {x[0], x[1]}};
      for (int k = 0; k < 3; k++) {
        const int64_t ls = m.localV(vp[k][0], vp[k][1]);
        if (ls >= 0) bad += (r.vij[k] != ls);
        else bad += check(KV, r.vij[k], m.ownerV(vp[k][0] + vp[k][1] * Nv), m.slotV1(vp[k][0], vp[k][1]));
      }
      if (bad) return fail("a record addresses a slot that does not hold its slice (tuple " + std::to_string(i) + " of batch " + std::to_string(ser) + ")");
    }
    return 0;
  };
  const int64_t per_call = (n + calls - 1) / calls;
  for (int64_t first = 0; first < n; first += per_call) {
    const int64_t count = std::min(per_call, n - first);
    serial += 2;  // run_list: everything of the previous call has completed
    const int64_t nb = (count + batch - 1) / batch;
    for (int64_t k = 0; k < nb; k++) {
      const Tuple *tp = t.data() + first + k * batch;
      const size_t nt = (size_t)std::min<int64_t>(batch, count - k * batch);
      plan_batch(m, tp, nt, owned, cache, serial + k, pl);
      if (pl.overflow) return fail("cache overflow");
      for (int q = 0; q < 3; q++) { stats[q] += (double)pl.used[q]; stats[3 + q] += (double)pl.hits[q]; }
      if (apply(pl, serial + k) || verify(pl, tp, nt, serial + k)) return -1;
      stats[7] += 1;
    }
    serial += nb;
    const int64_t rest = n - (first + count);
    if (rest > 0) {  // prefetch for the next call
      plan_batch(m, t.data() + first + count, (size_t)std::min(batch, rest), owned, cache, serial, pl);
      if (pl.overflow) cache.invalidate();
      else {
        for (int q = 0; q < 3; q++) stats[q] += (double)pl.used[q];
        if (apply(pl, serial)) return -1;
        serial += 1;
      }
    }
  }
  if (out8) std::memcpy(out8, stats, sizeof(stats));
  return 0;
}
int atrip_b200_host_cache_need(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                               int64_t batch, int64_t *out3) {
  if (nranks < 1 || Nv < 1 || rank < 0 || rank >= nranks || n < 0 || batch < 1) { g_error = "atrip_b200_host_cache_need: bad arguments"; return 1; }
  const ShardMap m(Nv, nranks, rank);
  std::vector<Tuple> t((size_t)n);
  for (int64_t i = 0; i < n; i++) t[i] = Tuple{abc[3 * i], abc[3 * i + 1], abc[3 * i + 2]};
  cache_need(m, t.data(), (size_t)n, (size_t)batch, out3);
  return 0;
}

int atrip_b200_host_store_source(int32_t store, int64_t No, int64_t Nv, int32_t a, int64_t x, int64_t y,
                                 int64_t row, int64_t kappa, double *out4) {
  if (No < 1 || Nv < 1 || x < 0 || x >= Nv || row < 0 || kappa < 0 || !out4 || (store != 0 && store != 1)) {
    g_error = "atrip_b200_host_store_source: bad arguments";
    return 1;
  }
  const SourceRef r = store == 0 ? ax_source_z((int)No, (int)Nv, a, (size_t)x, (size_t)row, (size_t)kappa)
                                 : by_source_z((int)No, (int)Nv, (size_t)x, (size_t)y, a, (size_t)row, (size_t)kappa);
  out4[0] = r.tensor;
  out4[1] = r.part;
  out4[2] = r.sign;
  out4[3] = (double)r.lin;
  return 0;
}
double atrip_b200_host_energy_z(int64_t No, double epsabc, const double *eps_i, const double *Tijk,
                                const double *Zijk, int32_t same) {
  return host_energy_z((int)No, epsabc, eps_i, Tijk, Zijk, same != 0);
}

}  // extern "C"
