/* atrip_b200.h -- C-ABI of the B200 device engine for atrip's (T) hot path.
 *
 * This is the boundary the host C++ (`atrip::Atrip::run`, include/atrip/Atrip.hpp,
 * atrip_b200/host/Atrip.cxx) talks through, and what any other host language would bind
 * (INTEGRATION.md).  Plain pointers and sizes only.  The reference has no C ABI of its own;
 * each entry point below names the reference code it replaces (paths relative to the
 * reference tree).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is available from
 *     atrip_b200_last_error() (thread-local).  The C++ host turns a non-zero status into a
 *     thrown std::string, which is what the reference throws (Acc.hpp:16-41).
 *   - all tensors are FP64 (real, or complex when atrip_b200_config.field = 1), column-major
 *     (first index fastest), exactly the layout CTF read_all / slice hands to the reference
 *     (SURVEY.md Appendix A.1).
 *   - one context drives one GPU; one process per GPU.  There is NO CPU fallback: every entry
 *     point that computes fails if the CUDA device is not usable.
 */
#ifndef ATRIP_B200_H
#define ATRIP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct atrip_b200_ctx atrip_b200_ctx;

typedef struct atrip_b200_config {
  int32_t device;       /* CUDA device ordinal (reference: rank % ngcards, Atrip.cxx:119) */
  int32_t rank;         /* this process' rank among the GPUs sharing the job */
  int32_t nranks;       /* number of GPUs / processes (GPU == "node" in group-and-sort) */
  int32_t with_J;       /* allocate the (cT) Jabci/Jijka stores (Atrip.cxx:338-362) */
  int64_t No, Nv;       /* occupied / virtual orbitals (lens[0] of epsilon_i/_a, Atrip.cxx:72-73) */
  int64_t batch_tuples; /* tuples per device batch; 0 = choose from No */
  int32_t resident;     /* 1: this rank stores every slice (replica); 0: only the slices it owns
                           (atrip_b200_host_slice_owner) plus a fetch cache filled over NCCL;
                           needs atrip_b200_comm_init before the first run when nranks > 1 */
  int32_t transport;    /* how remote slices travel when resident = 0 (ignored otherwise):
                           0 = default (currently 2), 1 = ncclSend/ncclRecv groups on a side stream,
                           the owner pushes what the peers' request lists ask for; 2 = peer-to-peer
                           pulls over NVLink by the copy engines (cudaMemcpyAsync from the owner's
                           store, mapped through CUDA IPC): no SM is taken from the contraction and
                           the owner does not take part */
  int32_t field;        /* 0: F = double (Atrip::run<double>, Atrip.cxx:1135); 1: F = std::complex<double>
                           (Atrip::run<Complex>, Atrip.cxx:1136).  With field = 1 EVERY tensor pointer of
                           this API (set_epsilon, set_Tai, load_*, tuple_debug cubes, read_slice) is an
                           array of interleaved (re, im) doubles = std::complex<double> memory layout,
                           still declared `double *`; energies stay real (Equations.cxx:176-178) */
} atrip_b200_config;

/* ---- lifecycle (replaces the ACC set-up in Atrip::run, Atrip.cxx:78-171, 217-218, 364-380) */
int atrip_b200_create(atrip_b200_ctx **ctx, const atrip_b200_config *cfg);
/* with sharded stores and transport 2, destroy is collective: it waits until every rank has
 * called it, because a peer may still be reading this rank's stores */
int atrip_b200_destroy(atrip_b200_ctx *ctx);
const char *atrip_b200_last_error(void);
const char *atrip_b200_version(void);
/* number of usable CUDA devices (0 if none; replaces cuDeviceGetCount, Atrip.cxx:82) */
int32_t atrip_b200_device_count(void);

/* ---- replicated small tensors (replaces read_all + HtoD, Atrip.cxx:176-215); host pointers */
int atrip_b200_set_epsilon(atrip_b200_ctx *ctx, const double *eps_i, const double *eps_a);
int atrip_b200_set_Tai(atrip_b200_ctx *ctx, const double *Tai /* [Nv,No] */);

/* ---- sliced tensors from HOST memory in CTF layout (replaces the SliceUnion constructors and
 *      slice_into_buffer: Unions.hpp:21-75, 77-278; SliceUnion.cxx:305-332).  The engine streams
 *      the tensor through pinned staging buffers and re-tiles it on the device into its own HBM
 *      layout (DESIGN.md "Data layout"); the host tensor may be freed afterwards
 *      (the reference deletes Vppph after slicing, Atrip.cxx:310).
 *      Order requirement: Tabij and Vijka before Vabci is NOT required; any order works. */
int atrip_b200_load_Tabij(atrip_b200_ctx *ctx, const double *Tabij /* [Nv,Nv,No,No] */);
int atrip_b200_load_Vabij(atrip_b200_ctx *ctx, const double *Vabij /* [Nv,Nv,No,No] */);
int atrip_b200_load_Vijka(atrip_b200_ctx *ctx, const double *Vijka /* [No,No,No,Nv] */);
int atrip_b200_load_Vabci(atrip_b200_ctx *ctx, const double *Vabci /* [Nv,Nv,Nv,No] */);
int atrip_b200_load_Jijka(atrip_b200_ctx *ctx, const double *Jijka /* [No,No,No,Nv] */);
int atrip_b200_load_Jabci(atrip_b200_ctx *ctx, const double *Jabci /* [Nv,Nv,Nv,No] */);

/* ---- sliced tensors, one slice at a time: what SliceUnion<F>::init does per rank (SliceUnion.cxx:305-332:
 *      for every source this rank owns, RankMap::find RankMap.cxx:35-85, CTF::slice it into a contiguous
 *      buffer, slice_into_vector Unions.hpp:21-75).  `host` holds n slices of one kind back to back in the
 *      REFERENCE's slice layout (column-major, what CTF::slice yields), xy their (x, y) indices (y ignored
 *      for the single-index kinds).  A rank uploads the slices atrip_b200_host_owned_slices lists for it
 *      and never needs the full tensors; slices the rank does not hold are skipped.  Kinds (Slice.hpp:99-108
 *      names):
 *        100 TA(x)      [Nv,No,No] = Tabij[x,:,:,:]   (TAPHH, Unions.hpp:77-113)
 *        101 VIJKA(x)   [No,No,No] = Vijka[:,:,:,x]   (HHHA,  Unions.hpp:115-152)      111 the same of Jijka
 *        200 VABCI(x,y) [Nv,No]    = Vabci[x,y,:,:]   (ABPH,  Unions.hpp:154-197)      210 the same of Jabci
 *        201 TABIJ(x,y) [No,No]    = Tabij[x,y,:,:], x <= y  (TABHH, Unions.hpp:239-278)
 *        202 VABIJ(x,y) [No,No]    = Vabij[x,y,:,:], x <= y  (ABHH,  Unions.hpp:199-237)
 *      Returns when the host buffer has been consumed.  With sharded stores and transport 2 the first
 *      upload after a run is collective (a peer may still be reading this rank's stores). */
int atrip_b200_upload_slices(atrip_b200_ctx *ctx, int32_t kind, int64_t n, const int64_t *xy /* n x 2 */,
                             const double *host);
int atrip_b200_upload_slice(atrip_b200_ctx *ctx, int32_t kind, int64_t x, int64_t y, const double *host);
/*      the inverse, for parity tests and for building host-side shards: n slices of one kind back to the
 *      reference layout; a slice this rank does not hold comes back as NaN */
int atrip_b200_read_slices(atrip_b200_ctx *ctx, int32_t kind, int64_t n, const int64_t *xy, double *out);
/*      host-only: the (x, y) list of the slices of `kind` rank `rank` of `nranks` has to be given; returns
 *      the count and writes at most cap pairs (xy may be NULL to query the count) */
int64_t atrip_b200_host_owned_slices(int32_t kind, int64_t Nv, int32_t rank, int32_t nranks, int64_t *xy,
                                     int64_t cap);

/* ---- synthetic inputs generated on the device from the counter-based generator
 *      value = f(seed, tensor id, column-major linear index) (DESIGN.md "Synthetic inputs";
 *      plays the role of CTF fill_random in bench/main.cxx:54-89, with ranges that keep the
 *      energy denominators away from zero).  Fills every tensor, epsilons and Tai included. */
int atrip_b200_fill_synthetic(atrip_b200_ctx *ctx, uint64_t seed, double scale);

/* ---- tuples (replaces TuplesDistribution::get_tuples, Tuples.cxx:136-141, 310-407)
 *      distribution: 0 = NAIVE order (all tuples, lexicographic, cut in contiguous chunks),
 *                    1 = GROUP_AND_SORT with one GPU per "node" (Tuples.cxx:156-308).
 *      The list of this rank is padded with FAKE_TUPLE {0,0,0} (Tuples.hpp:43) to the longest
 *      rank's length, as the reference does (Tuples.cxx:346-377). */
int atrip_b200_build_tuples(atrip_b200_ctx *ctx, int32_t distribution);
int atrip_b200_set_tuples(atrip_b200_ctx *ctx, const uint64_t *abc /* n x 3 */, int64_t n);
int64_t atrip_b200_num_tuples(const atrip_b200_ctx *ctx);
int atrip_b200_get_tuples(const atrip_b200_ctx *ctx, uint64_t *abc /* n x 3 */, int64_t cap);

/* ---- execute (replaces the main loop body, Atrip.cxx:686-1057: doubles_contribution,
 *      singles_contribution, get_energy_distinct/same per tuple).  Runs tuples
 *      [first, first+count) of this rank's list and returns the partial sums
 *      sum_t e_abc (NOT yet negated; the caller all-reduces and negates, Atrip.cxx:1094-1111).
 *      ct_energy may be NULL.  Blocking; the device work is asynchronous inside. */
int atrip_b200_run(atrip_b200_ctx *ctx, int64_t first, int64_t count, double *energy,
                   double *ct_energy);

/* ---- debug / parity: one tuple, returning the reference's Tijk and Zijk cubes
 *      (No^3 each, [i + j No + k No^2]; either may be NULL) and its energy contribution */
int atrip_b200_tuple_debug(atrip_b200_ctx *ctx, int64_t a, int64_t b, int64_t c, double *Tijk,
                           double *Zijk, double *energy);

/* ---- read back one slice in the REFERENCE's slice layout, for parity of the ingest/fill
 *      paths.  kind: 100 TA(x) [Nv,No,No]; 101 VIJKA(x) [No,No,No]; 200 VABCI(x,y) [Nv,No];
 *      201 TABIJ(x,y) [No,No]; 202 VABIJ(x,y) [No,No]  (Slice.hpp:99-108 names) */
int atrip_b200_read_slice(atrip_b200_ctx *ctx, int32_t kind, int64_t x, int64_t y, double *out);

/* ---- communicator of the job (replaces Atrip::init(MPI_Comm), Atrip.cxx:54-63, and the MPI calls
 *      on the hot path: MPI_Isend/Irecv of slices, SliceUnion.cxx:456-462, 491-503, and the final
 *      MPI_Reduce, Atrip.cxx:1094-1107).  One NCCL communicator with one rank per GPU: rank 0
 *      calls atrip_b200_comm_unique_id and ships the 128 bytes to the other ranks through
 *      whatever the host has (MPI_Bcast in the C++ API, torch.distributed in bench.py), then every
 *      rank calls atrip_b200_comm_init (collective; with transport 2 it also maps the peers'
 *      stores).  With sharded stores and transport 1 atrip_b200_run and atrip_b200_tuple_debug are
 *      COLLECTIVE: every rank calls them with the same count (lists are padded to equal length
 *      with the fake tuple for exactly this reason).  With transport 2 only the first run after
 *      the stores were (re)filled synchronises the ranks. */
int atrip_b200_comm_unique_id(void *id128);
int atrip_b200_comm_init(atrip_b200_ctx *ctx, const void *id128);
/*      in-place SUM over all ranks of n <= 16 doubles in host memory (ncclAllReduce) */
int atrip_b200_allreduce(atrip_b200_ctx *ctx, double *vals, int32_t n);
/*      slice traffic of the last run on this rank: out[0] = bytes received, out[1] = messages */
int atrip_b200_last_exchange(const atrip_b200_ctx *ctx, double *out2);

/* ---- debug: order-independent checksum (integer sum of the bit patterns) of the class cubes the
 *      contraction kernel wrote for the LAST batch of the last run -- separates "contraction output
 *      changed" from "reduction changed" when chasing run-to-run differences */
int atrip_b200_debug_cubes_checksum(atrip_b200_ctx *ctx, uint64_t *out);

/* ---- timing of the last atrip_b200_run, measured with CUDA events on the engine's stream:
 *      out[0] = total ms, out[1] = contraction kernel ms, out[2] = reduction kernel ms,
 *      out[3] = number of contraction launches, out[4] = number of reduction launches,
 *      out[5] = non-fake tuples processed */
int atrip_b200_last_timing(const atrip_b200_ctx *ctx, double *out6);

/*      phases of the last atrip_b200_run: out[0] = ms the compute stream sat idle in front of contraction
 *      launches of batches 1.. (slice fetch not landed / cube buffer not yet reduced; device events),
 *      out[1] = host ms spent building slot records and fetch schedules (replaces build_local_database,
 *      SliceUnion.cxx:36-171), out[2] = remote slices found in the fetch cache (the reference's Recycled /
 *      exact-match cases, SliceUnion.cxx:66-137), out[3] = remote slices fetched, out[4] = batches,
 *      out[5] = ms of idle gap in front of batch 0 (start-up of the call) */
int atrip_b200_last_phases(const atrip_b200_ctx *ctx, double *out6);

/* ---- derived constants the caller needs for reporting */
int64_t atrip_b200_kp(const atrip_b200_ctx *ctx);            /* padded contraction length */
double atrip_b200_flops_per_tuple(const atrip_b200_ctx *ctx); /* 12 No^3 (No+Nv), x4 complex; Atrip.cxx:578-580 */
int64_t atrip_b200_batch_tuples(const atrip_b200_ctx *ctx);  /* tuples per contraction launch */

/* ---- measurement helpers
 *      FP64 tensor-core ceiling of the device, measured live with a register-resident
 *      DMMA.8x8x4 loop (about 0.2 s); the roofline denominator of the contraction kernel
 *      (MEASURED_PEAKS.json carries no FP64 figure).  Returns TFLOP/s in *tflops. */
int atrip_b200_measure_dmma_peak(int32_t device, double *tflops);
/*      counter-based synthetic values of one input tensor in CTF (column-major) order,
 *      elements [first, first+count), generated on the device and copied to host memory;
 *      lets a caller build host tensors of bench size without a multi-minute CPU fill.
 *      tensor_id: 0 eps_i, 1 eps_a, 2 Tai, 3 Tabij, 4 Vabij, 5 Vijka, 6 Vabci, 7 Jijka, 8 Jabci */
int atrip_b200_synth_to_host(int32_t device, uint64_t seed, int32_t tensor_id, double scale, uint64_t first,
                             uint64_t count, double *host);

/*      contraction-kernel plan chosen for No (host-only, DESIGN.md "Tile planner"):
 *      out[0..10] = MI, NI, consumer warps, tu, tv, row tiles, column tiles, pipeline stages,
 *      dynamic shared memory bytes, A rows per stage, useful-DMMA fraction x 1e6 */
int atrip_b200_host_plan(int64_t No, int64_t smem_limit_bytes, int64_t *out11);

/* ---- host-only utilities (no device needed; usable before any context exists)
 *      tuple list of rank `rank` of `nranks` (3 x uint64 per tuple, padded with the fake tuple
 *      when pad != 0); returns the list length, writes at most cap tuples.
 *      distribution as in atrip_b200_build_tuples. */
int64_t atrip_b200_host_tuples(int32_t distribution, int64_t Nv, int32_t rank, int32_t nranks, int32_t pad,
                               uint64_t *abc, int64_t cap);
/*      owner rank of a slice (replaces RankMap<F>::find, RankMap.cxx:35-85, one rank per node):
 *      kind 100/101 -> x % nranks; pair kinds (x,y) -> x % nranks: a pair lives with its FIRST
 *      index, which equals the reference's (x + y Nv) % nranks (RankMap.cxx:43-44) whenever
 *      Nv % nranks == 0 and keeps the co-location for any Nv.  VABIJ(x<=y) is additionally
 *      replicated on y % nranks.  *slot (may be NULL) receives the slot in the owner's store. */
int32_t atrip_b200_host_slice_owner(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t nranks);
int32_t atrip_b200_host_slice_slot(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t nranks, int64_t *slot);
/*      slot of a slice in the store of `rank`, or -1 if that rank does not hold it; kind 203 =
 *      the transposed-hole twin (x,x)' of the diagonal pair slice */
int64_t atrip_b200_host_local_slot(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t rank, int32_t nranks);
/*      owned slots per store of `rank`: out[0] AX (TAPHH+HHHA), out[1] BY (ABPH+TABHH, ordered
 *      pairs + transposed diagonal), out[2] VIJ (ABHH) */
int atrip_b200_host_shard_sizes(int64_t Nv, int32_t rank, int32_t nranks, int64_t *out3);
/*      fetch schedule of one device batch (replaces build_local_database, SliceUnion.cxx:36-171,
 *      and the per-tuple database exchange, Atrip.cxx:414-459): for the n tuples abc of `rank`
 *      writes recs (n x 16 int32: a b c fake, AX slots of a b c, BY slots of (b,c) (a,c) (c,b)'
 *      (a,b) (c,a)' (b,a)', VIJ slots of (b,c) (a,c) (a,b); slots >= owned count address the cache,
 *      cache_base3 = slot number of cache slot 0 per store = owned count) and up to cap fetch ranges
 *      (5 x int64 each: peer, store 0/1/2, first slot at the owner, count, first cache slot;
 *      planned against an EMPTY fetch cache, so cache slots count up from 0).  Returns the number of
 *      ranges, < 0 on error. */
int64_t atrip_b200_host_plan_batch(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                                   const int64_t *cache_base3, int32_t *recs, int64_t *ranges, int64_t cap);
/*      the whole list cut into `calls` run calls and walked batch by batch through the persistent fetch
 *      cache (cap3 slots per store) exactly as atrip_b200_run does, checked against an independent model
 *      of the cache contents: every record addresses a slot that holds its slice, no copy overwrites a
 *      slot of a batch that may still compute.  out[0..2] slices fetched per store, out[3..5] cache hits,
 *      out[6] copy ranges, out[7] batches.  Returns 0, or -1 with the violation in last_error. */
int atrip_b200_host_check_schedule(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                                   int64_t batch, int32_t calls, const int64_t *cap3, double *out8);
/*      cache slots per store that any window of `batch` consecutive tuples of the list needs */
int atrip_b200_host_cache_need(int64_t Nv, int32_t rank, int32_t nranks, const uint64_t *abc, int64_t n,
                               int64_t batch, int64_t *out3);

/* ---- complex field, host-only checks of the device code's shared __host__ __device__ pieces
 *      (CPU test-suite hooks; no device needed).
 *      atrip_b200_host_store_source: which source-tensor element a store element of the complex
 *      layout holds (stores.cuh "complex field").  store 0 = AX (a = variant, x = virtual index,
 *      row = p + q No), store 1 = BY (a = transposed-diagonal flag, x,y = ordered pair, row = r).
 *      out[0] = tensor (0 padding, 1 Tabij, 2 Vijka, 3 Vabci), out[1] = part (0 re, 1 im),
 *      out[2] = sign, out[3] = column-major element index in that tensor.
 *      atrip_b200_host_energy_z: get_energy_distinct/same<Complex> (Equations.cxx:101-238) over plain
 *      interleaved-complex No^3 cubes through the per-point function the device kernel uses. */
int atrip_b200_host_store_source(int32_t store, int64_t No, int64_t Nv, int32_t a, int64_t x, int64_t y,
                                 int64_t row, int64_t kappa, double *out4);
double atrip_b200_host_energy_z(int64_t No, double epsabc, const double *eps_i, const double *Tijk,
                                const double *Zijk, int32_t same);

#ifdef __cplusplus
}
#endif
#endif
