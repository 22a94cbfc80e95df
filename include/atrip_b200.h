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
 *   - all tensors are FP64, column-major (first index fastest), exactly the layout CTF
 *     read_all / slice hands to the reference.
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
                           by RankMap round-robin plus a fetch cache */
  int32_t reserved;
} atrip_b200_config;

/* ---- lifecycle (replaces the ACC set-up in Atrip::run, Atrip.cxx:78-171, 217-218, 364-380) */
int atrip_b200_create(atrip_b200_ctx **ctx, const atrip_b200_config *cfg);
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

/* ---- timing of the last atrip_b200_run, measured with CUDA events on the engine's stream:
 *      out[0] = total ms, out[1] = contraction kernel ms, out[2] = reduction kernel ms,
 *      out[3] = number of contraction launches, out[4] = number of reduction launches,
 *      out[5] = non-fake tuples processed */
int atrip_b200_last_timing(const atrip_b200_ctx *ctx, double *out6);

/* ---- derived constants the caller needs for reporting */
int64_t atrip_b200_kp(const atrip_b200_ctx *ctx);            /* padded contraction length */
double atrip_b200_flops_per_tuple(const atrip_b200_ctx *ctx); /* 12 No^3 (No+Nv), Atrip.cxx:578-580 */
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
 *      kind 100/101 -> x % nranks; pair kinds -> (x + y Nv) % nranks (RankMap.cxx:43-44) */
int32_t atrip_b200_host_slice_owner(int32_t kind, int64_t x, int64_t y, int64_t Nv, int32_t nranks);

#ifdef __cplusplus
}
#endif
#endif
