/*
 * nemb200.h -- C ABI of the B200-native NEM partitioning path (libnemb200.so, sm_100a only).
 *
 * Drop-in boundary.  In the reference the only native entry point on this path is
 *     int nem(const char* Fname, int nk, const char* algo, float beta, const char* convergence, float convergence_th,
 *             const char* format, int it_max, int dolog, const char* model_family, const char* proportion,
 *             const char* dispersion, int init_mode, const char* init_file, const char* out_file_prefix, int seed)
 * (/root/reference/ppanggolin/nem/NEM/nem_exe.h:23-38, bound by NEM/nem_stats.pyx:1-17, called from
 * ppanggolin/nem/partition.py:126-150) and every operand travels through text files (partition.py:344-436 writes
 * .str/.dat/.nei/.index, :65-85 writes the init .m, :170-231 parses .uf/.mf).  The entry points below replace that
 * call with the same operands passed in memory:
 *     .dat  -> `bits`   bit-packed presence matrix (families x genomes)
 *     .nei  -> `rowptr/col/w`  CSR neighbour lists with contiguity weights, as nem_exe.c:1347-1483 keeps them
 *     .m    -> built on the device from K and D (partition.py:65-85 is a pure function of K, D)
 *     .uf   -> `T`      posteriors (float32, N x K);  .mf -> `crit`, `mu`, `eps`, `pk`
 * nem()'s constant arguments in PPanGGOLiN (algo "nem", convergence "clas", format "fuzzy", family "bern",
 * proportion "pk", init_mode 2) are the only supported configuration and are therefore not parameters;
 * dispersion "sk_"/"skd" is `free_disp`; `seed` has no effect on this path.
 *
 * Conventions: plain pointers and sizes, no torch types.  Matrix layout: row-major uint32 words, bit b of word w of a
 * row is genome column 32*w+b; `row_stride_words` is a multiple of 4 (rows start 16-byte aligned) and padding bits
 * are zero; for wide matrices a multiple of 32 (whole 128-byte lines) is what the streaming kernels run fastest on.  All functions return NEMB200_OK (0) or a negative error; nemb200_last_error() gives the message of the
 * calling thread's last failure.  A run whose EM hits an empty class reports it in result->status (the reference
 * then writes no .uf and partition.py:233-265 returns ({}, None, None)); that is not an API error.
 * There is no CPU fallback: without a CUDA device every compute entry fails with NEMB200_ERR_CUDA.
 */
#ifndef NEMB200_H
#define NEMB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEMB200_ABI_VERSION 2
#define NEMB200_MAX_K 32

enum { NEMB200_OK = 0, NEMB200_ERR_ARG = -1, NEMB200_ERR_CUDA = -2, NEMB200_ERR_NOMEM = -3 };

/* result->status: what the reference's StatusET would have been (nem_exe.c:642-708) */
enum { NEMB200_STS_OK = 0, NEMB200_STS_EMPTYCLASS = 1, NEMB200_STS_INFINITE = 2 };

/* flags */
enum {
  NEMB200_EPI_REF      = 0,      /* E-step epilogue with the reference's arithmetic: float32 log-density, double
                                    exp() that may underflow, S==0 -> 1/K (nem_alg.c:2660-2698)                  */
  NEMB200_EPI_LOGSPACE = 1 << 0, /* log-sum-exp epilogue: valid for any number of genomes */
  NEMB200_JACOBI       = 1 << 1, /* parallel posterior update instead of the reference's in-place Gauss-Seidel
                                    sweep (nem_alg.c:2459-2464); off by default                                  */
  NEMB200_PEER_EXCHANGE = 1 << 3, /* plan owns a full-length log-density buffer and signal words that other ranks of a
                                    row-sharded run map through CUDA IPC (nemb200_peer_*)                        */
  NEMB200_FULL_RECOUNT = 1 << 2, /* M-step: recount every one-hot row at every iteration instead of updating the integer
                                    column counts with the rows that changed class (same result bit for bit)    */
  NEMB200_TIME_DENSITY = 1 << 4  /* whole-run entries: bracket every density launch of an un-batched (large) instance
                                    with CUDA events and report their sum in result->density_ms (the benchmark's
                                    roofline figure of the dominant kernel, measured inside the timed runs)       */
};

typedef struct nemb200_result {
  double  crit[5];    /* U, D, L, M, Z  (.mf line 3, nem_exe.c:1713-1717); partition.py:187 reads crit[3] */
  int32_t n_iter;     /* EM iterations executed (".stderr": "NEM converged after n iterations") */
  int32_t converged;
  int32_t status;     /* NEMB200_STS_* */
  int32_t n_levels;   /* depth of the Gauss-Seidel level schedule that was used */
  float   elapsed_ms; /* device time of the run (CUDA events): start phase, EM iterations and criteria, excluding copies;
                         for an instance of a batch: of its whole group */
  float   em_ms;      /* of which: from the start phase to the end of the last EM iteration (no criteria) */
  float   density_ms; /* NEMB200_TIME_DENSITY: device time of the density launches of the run's EM iterations (CUDA */
  int32_t density_launches; /* events around each launch, on the launching stream) and their number; else 0 */
} nemb200_result;

int         nemb200_abi_version(void);
const char* nemb200_last_error(void);
int         nemb200_device_count(void);
/* Plans and whole-run calls take their device memory from a per-process cache of freed blocks (cudaMalloc/cudaFree are
 * too slow for thousands of chunk instances); this returns the cache to the driver. */
int         nemb200_release_cached_memory(void);

/* ---- whole-run entry points (replace nem(), nem_exe.h:23-38) ------------------------------------------------- */

/* Host buffers in, host buffers out; host<->device copies happen inside.  T_out: N*K, mu_out/eps_out: K*D,
 * pk_out: K (any may be NULL). */
int nemb200_nem_host(const uint32_t* bits, int64_t N, int64_t D, int64_t row_stride_words,
                     const int32_t* rowptr, const int32_t* col, const float* w,
                     int32_t K, float beta, int32_t free_disp, int32_t itermax, float cv_th, uint32_t flags,
                     float* T_out, float* mu_out, float* eps_out, float* pk_out, nemb200_result* result);

/* Same with the matrix, CSR and outputs resident in device memory (device pointers); runs on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream).  rowptr/col/w may be NULL when beta == 0.
 * `sched_*`: Gauss-Seidel level schedule in device memory from nemb200_gs_schedule (may be NULL when beta == 0 or
 * NEMB200_JACOBI is set). */
int nemb200_nem_device(const uint32_t* bits, int64_t N, int64_t D, int64_t row_stride_words,
                       const int32_t* rowptr, const int32_t* col, const float* w,
                       const int32_t* sched_level_ptr, const int32_t* sched_rows, int32_t n_levels,
                       int32_t K, float beta, int32_t free_disp, int32_t itermax, float cv_th, uint32_t flags,
                       float* T_out, float* mu_out, float* eps_out, float* pk_out, nemb200_result* result,
                       void* stream);

/* A batch of independent instances (the K-selection sweep of partition.py:482-514: same matrix, K = 2..20, beta 0,
 * 10 iterations; or the chunk samples of partition.py:837-846), device-resident operands, arrays of length n_inst.
 * Every instance gets its own stream; iterations are enqueued in lock-step so that the small kernels of different
 * instances overlap on the GPU, and there is one flag read per instance and iteration.  Array arguments that are not
 * needed (graph arrays when every beta is 0, output arrays) may be NULL. */
int nemb200_nem_device_batch(int32_t n_inst, const uint32_t* const* bits, const int64_t* N, const int64_t* D,
                             const int64_t* row_stride_words, const int32_t* const* rowptr, const int32_t* const* col,
                             const float* const* w, const int32_t* const* sched_level_ptr,
                             const int32_t* const* sched_rows, const int32_t* n_levels, const int32_t* K,
                             const float* beta, const int32_t* free_disp, const int32_t* itermax, float cv_th,
                             uint32_t flags, float* const* T_out, float* const* mu_out, float* const* eps_out,
                             float* const* pk_out, nemb200_result* results);

/* Host helper: level schedule of the in-place sweep.  level(i) = 1 + max level(j) over neighbours j < i, so rows of one
 * level can be updated concurrently and still read exactly the values the sequential sweep 0..N-1 would (neighbours
 * j < i already updated, j >= i not yet; nem_alg.c:2441-2476).  level_ptr_out: N+1 ints (only n_levels+1 used),
 * rows_out: N ints.  Returns n_levels (>= 1) or a negative error. */
int nemb200_gs_schedule(int64_t N, const int32_t* rowptr, const int32_t* col, int32_t* level_ptr_out, int32_t* rows_out);

/* ---- stepwise entry points (device pointers): used by the row-sharded multi-GPU driver and by the parity tests -- */

typedef struct nemb200_plan nemb200_plan; /* opaque: device workspace for one (N rows on this rank, D, K) shape */

int  nemb200_plan_create(nemb200_plan** plan, int64_t N_local, int64_t N_total, int64_t D, int64_t row_stride_words,
                         int32_t K, int32_t free_disp, uint32_t flags);
void nemb200_plan_destroy(nemb200_plan* plan);

/* Measurement aid: while enabled, every k_mstep_onehot launch is bracketed by CUDA events on its stream.  Each call
 * returns (and clears) the time and count accumulated so far, then sets the mode.  Synchronises on the recorded events. */
int nemb200_plan_timing(nemb200_plan* plan, int32_t enable, float* onehot_ms_total, int32_t* onehot_launches);

/* partition.py:65-85 + nem_exe.c:1023-1040: initial pk/mu/eps for (K, D), written into the plan's parameter block. */
int nemb200_init_params(nemb200_plan* plan, void* stream);

/* M-step, part 1 (nem_mod.c:1270-1312, :1641-1699 restated as column sums): accumulate this rank's sufficient
 * statistics from posteriors T (N_local x K, device).  Outputs live in the plan; nemb200_stats_ptrs exposes them so a
 * multi-GPU driver can all-reduce them in place: cnt = int32[K][Dpad] one-hot column counts,
 * fsum = double[K][Dpad] fuzzy column sums, nk_cnt = int64-free int32[K], nk_fz = double[K]. */
int nemb200_mstep_accumulate(nemb200_plan* plan, const uint32_t* bits, const float* T, void* stream);
int nemb200_stats_ptrs(nemb200_plan* plan, int32_t** cnt, double** fsum, int32_t** nk_cnt, double** nk_fz, int64_t* Dpad);
/* The same statistics as two contiguous blocks (cnt followed by nk_cnt; fsum followed by nk_fz), so that one int32 and
 * one fp64 all-reduce cover them. */
int nemb200_stats_blocks(nemb200_plan* plan, int32_t** iblock, int64_t* ilen, double** dblock, int64_t* dlen);

/* M-step, part 2 (nem_mod.c:1317-1474 median, :1020-1076 / :1131-1170 dispersion, :455-459 proportions): parameters
 * from (all-reduced) statistics.  Sets the plan's status word to NEMB200_STS_EMPTYCLASS when a class is empty. */
int nemb200_mstep_finalize(nemb200_plan* plan, void* stream);

/* E-step, part 1 (nem_mod.c:618-689 + nem_alg.c:2319-2374): log f_k(x_i) for this rank's rows -> logf (N_local x K
 * doubles, device; -inf encodes the reference's null density). */
int nemb200_density(nemb200_plan* plan, const uint32_t* bits, double* logf, void* stream);

/* ---- row-sharded runs: the two exchange steps of an iteration over NVLink peer memory, fused into the kernels ------
 * One process per GPU; every rank creates its plan with NEMB200_PEER_EXCHANGE, exports (CUDA IPC handle of the plan's
 * block + 16 offsets), the driver all-gathers those 192 bytes per rank out of band (torch.distributed), and every rank
 * attaches.  From then on
 *   - nemb200_density_peers stores the shard's log-densities into the log-density buffer of EVERY rank (peer stores)
 *     and publishes an epoch; nemb200_peer_wait_densities blocks the stream until all ranks have published it: together
 *     they are the all-gather that precedes the replicated Gauss-Seidel sweep;
 *   - nemb200_mstep_finalize publishes / waits for the statistics epoch and reads every rank's statistic blocks (peer
 *     loads), adding them in rank order: the all-reduce of the M-step, identical parameters on every rank.
 * world <= 8 (one NVSwitch box).  The reference has no counterpart (one process per NEM instance, partition.py:505). */
#define NEMB200_PEER_OFFSETS 16
int nemb200_peer_export(nemb200_plan* plan, unsigned char handle_out[64], int64_t offsets_out[NEMB200_PEER_OFFSETS]);
int nemb200_peer_attach(nemb200_plan* plan, int32_t world, int32_t rank, const unsigned char* handles /*world*64*/,
                        const int64_t* offsets /*world*NEMB200_PEER_OFFSETS*/);
int nemb200_peer_logf(nemb200_plan* plan, double** logf_full);
int nemb200_density_peers(nemb200_plan* plan, const uint32_t* bits, int64_t row_offset, void* stream);
int nemb200_peer_wait_densities(nemb200_plan* plan, void* stream);

/* A whole row-sharded run with the device-resident loop (no host synchronisation between iterations): this rank's part.
 * `plan` = created with NEMB200_PEER_EXCHANGE for this rank's rows [row_offset, row_offset + N_local) and attached; every
 * rank calls it with the same graph (all rows), beta and iteration cap.  The convergence test runs on the replicated
 * posteriors, so all ranks stop at the same iteration.  T_out: N_total x K (device, may be NULL); the parameters stay in
 * the plan (nemb200_get_params). */
int nemb200_nem_sharded(nemb200_plan* plan, const uint32_t* bits_local, int64_t row_offset, const int32_t* rowptr,
                        const int32_t* col, const float* w, const int32_t* sched_level_ptr, const int32_t* sched_rows,
                        int32_t n_levels, float beta, int32_t itermax, float cv_th, float* T_out, nemb200_result* result,
                        void* stream);

/* E-step, part 2 (nem_alg.c:2412-2489, :2630-2701): posterior sweep over rows [0, N) of logf/T (full-length arrays),
 * reading T_old and writing T_new; maxdiff_bits receives max |T_new - T_old| as float bits (atomicMax).
 * beta_now lets the caller run the "blind" beta = 0 pass of nem_alg.c:2043-2050.  maxdiff_bits == NULL uses (and first
 * resets) the plan's own convergence word, which nemb200_read_flags returns together with the status word. */
int nemb200_sweep(nemb200_plan* plan, int64_t N, const double* logf, const float* T_old, float* T_new,
                  const int32_t* rowptr, const int32_t* col, const float* w,
                  const int32_t* sched_level_ptr, const int32_t* sched_rows, int32_t n_levels,
                  float beta_now, uint32_t* maxdiff_bits, void* stream);

/* A level-scheduled sweep keeps a schedule-ordered copy of the neighbour graph inside the plan, built on first use and
 * reused while the same rowptr / col / w / sched_rows POINTERS and N are passed.  A caller that changes those arrays in
 * place, or reuses their addresses for another graph, calls this before the next nemb200_sweep. */
int nemb200_plan_graph_changed(nemb200_plan* plan);

/* Diagnostics.  The criteria D and G are float32 sums taken in the reference's order (nem_alg.c:2813-2823); large
 * instances compute them with a parallel scheme that must give the bits of the sequential sum.  This entry runs either
 * implementation on caller-supplied terms (host arrays of n floats each) so that tests can compare them on adversarial
 * inputs: out2[0] = sum of d, out2[1] = sum of g, each accumulated in float32 from 0 in index order. */
int nemb200_selftest_float_chain(const float* d, const float* g, int64_t n, int32_t parallel, double out2[2]);
/* The criteria L and Z (nem_alg.c:2826-2827) are float32 accumulators that receive DOUBLE terms, acc = (float)((double)acc
 * + x): the same two implementations for that kind of chain, on caller-supplied double terms. */
int nemb200_selftest_double_chain(const double* l, const double* z, int64_t n, int32_t parallel, double out2[2]);

/* nem_alg.c:2160-2173 (HasConverged, CVTEST_CLAS) needs max |T_new - T_old|; nem_alg.c:1873-1892 needs the empty-class
 * status.  Both in one 8-byte device->host read (synchronises the stream). */
int nemb200_read_flags(nemb200_plan* plan, float* maxdiff, int32_t* status, void* stream);

/* nem_alg.c:2764-2843: criteria U, D, L, M, Z for final posteriors T (host output, synchronises the stream). */
int nemb200_criteria(nemb200_plan* plan, int64_t N, const double* logf, const float* T,
                     const int32_t* rowptr, const int32_t* col, const float* w, float beta, double crit_out[5],
                     void* stream);

/* Copy the plan's current parameters / status to host buffers (mu, eps: K*D; pk: K; any may be NULL). */
int nemb200_get_params(nemb200_plan* plan, float* mu, float* eps, float* pk, int32_t* status, void* stream);

/* partition.py:206-231 on the "%5.3f" text of nem_exe.c:1682, on the device: labels[i] = class 0..K-1, or -1 for "S_"
 * (tie or max < 0.5); *entropy = sum p ln p over the 3-decimal posteriors (partition.py:212-220).  T: device. */
int nemb200_labels(const float* T, int64_t N, int32_t K, int32_t* labels_dev, double* entropy_host, void* stream);

/* ---- exporter: NEM instances derived on the device from the array form of a pangenome ---------------------------------
 * Replaces write_nem_input_files (ppanggolin/nem/partition.py:344-436), which walks the Python objects again for every
 * genome subset (the whole pangenome, the K-selection sample, every chunk of partition.py:809-821) and writes text.
 * The pangenome is described ONCE by arrays in device memory; a call derives the instances of a batch of genome subsets:
 * matrix rows of the families present in the subset (:381-391), neighbour lists weighted round(gene pairs in the subset /
 * #genomes, 4) (:393-414) with the max-degree cut (:415-433) exactly as the C reader keeps them (NEM/nem_exe.c:1395-1451),
 * sum of kept weights / 2 (:416,:436), and the Gauss-Seidel level schedule (nemb200_gs_schedule) -- without a host round
 * trip of the arrays. */
typedef struct nemb200_pangenome {
  int64_t F, G, E, A;           /* families, genomes, edges, adjacency entries */
  int64_t Wg;                   /* words per presence row, >= ceil(G / 32) */
  const uint32_t* presence;     /* [F][Wg] bit g of row f: family f has a gene in genome g (geneFamily.py:296) */
  const int32_t*  adj_ptr;      /* [F+1] adjacency rows, entries in fam.edges order (geneFamily.py:272) */
  const int32_t*  adj_other;    /* [A]   the other family of the edge */
  const int32_t*  adj_edge;     /* [A]   edge id */
  const int32_t*  cov_ptr;      /* [E+1] gene pairs per (edge, genome) as CSR over edges (edge.py:75 get_organisms_dict), or NULL: */
  const int32_t*  cov_org;      /* [M]   genome                                                                                   */
  const int32_t*  cov_cnt;      /* [M]   number of gene pairs                                                                     */
  const int32_t*  edge_fam;     /* [E][2] with cov_ptr == NULL: one gene pair in genome g iff both families are present in g
                                   (synthetic pangenomes at scales where the explicit table does not fit)                        */
} nemb200_pangenome;

typedef struct nemb200_sample {
  const int32_t* sel;           /* in  (device) genome columns of the subset, ascending: column j of the instance;
                                   NULL = all genomes in order (then n_sel = G and row_stride_words = Wg)          */
  int32_t  n_sel;
  int64_t  row_stride_words;    /* in: words per instance row, multiple of 4 */
  uint32_t* bits;               /* out (device, capacity F rows): matrix of the instance */
  int32_t* fam_rows;            /* out (device, F): family index of every instance row */
  int32_t* rowptr;              /* out (device, F + 1) */
  int32_t* col;                 /* out (device, A) */
  float*   w;                   /* out (device, A) */
  int32_t* level_ptr;           /* out (device, F + 1): level schedule, may be NULL with want_levels == 0 */
  int32_t* level_rows;          /* out (device, F) */
  int64_t  N, nnz;              /* out (host): rows and neighbour entries of the instance */
  int32_t  n_levels;
  double   edges_weight;        /* out (host): partition.py:436, the caller scales beta with it (:324, :872) */
} nemb200_sample;

/* weight_lut (device, lut_len floats): lut[c] = (float) round(c / n_sel, 4) as Python computes it (partition.py:413), for
 * every pair count c that can occur; all samples of a call share n_sel.  Synchronises `stream`. */
int nemb200_export_samples(const nemb200_pangenome* pan, int32_t n_samples, nemb200_sample* samples, float sm_degree,
                           const float* weight_lut, int32_t lut_len, int32_t want_levels, void* stream);

/* validate_family (partition.py:784-802) on the device, in two steps.  nemb200_chunk_codes, per finished chunk: the label
 * rule of partition.py:206-231 on the posteriors T (N x K) reduced to the letter that is voted with (0 = P, 1 = S, 2 = C),
 * scattered into codes[F] (int8, filled with -1 by the caller: family not in the chunk).  nemb200_apply_votes, per round:
 * every family applies the votes of the round's chunks in sample order (codes[n_chunks][F]) to votes[F][4] (P, S, C, U);
 * families that become validated are marked in validated[F] and counted in *n_validated (device int).
 * ratio = n_org / chunk_size. */
int nemb200_chunk_codes(const float* T, int64_t N, int32_t K, const int32_t* fam_rows, int8_t* codes, void* stream);
int nemb200_apply_votes(const int8_t* codes, int32_t n_chunks, int64_t F, int32_t* votes, uint8_t* validated,
                        int32_t* n_validated, double ratio, int32_t n_org, void* stream);

/* rarefaction.py:657-684 (core / soft-core counts of a genome sample, gmpy2 bitarrays there): counts[f] = number of the
 * selected genomes that hold family f.  mask_scratch: Wg words (device). */
int nemb200_subset_counts(const nemb200_pangenome* pan, const int32_t* sel, int32_t n_sel, uint32_t* mask_scratch,
                          int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEMB200_H */
