/* rqb200.h -- C ABI of the B200 hot path underneath the nanorq.h API.
 *
 * These are the entry points a reference-side binding would call instead of
 * the reference's internal solver seam:
 *
 *   rqb_solver_*      replaces precode_matrix_gen / precode_matrix_invert /
 *                     precode_matrix_intermediate (reference include/precode.h:10-12,
 *                     lib/precode.c:90,347,379) and decode_row (lib/nanorq.c:184)
 *                     for one source block resident on the GPU;
 *   rqb_rowops_*      replaces the per-row oblas seam oaxpy / oaddrow / oscal
 *                     (reference deps/oblas/oblas.h:28-30) with a batched call;
 *   rqb_schedule_*    replays a reference-format schedule (sched_op list + marks,
 *                     include/sched.h:6-27; lib/precode.c:23-32) on the device.
 *
 * Plain pointers and sizes only.  All functions return 0 on success; negative
 * values are argument/state errors, positive values are documented per call.
 * Nothing here falls back to the CPU: without a CUDA device every compute
 * entry point fails with RQB_E_NODEVICE.
 */
#ifndef RQB200_H
#define RQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RQB_OK 0
#define RQB_NEED_MORE 1      /* matrix rank < L: add symbols and retry     */
#define RQB_E_ARG (-1)
#define RQB_E_NODEVICE (-100) /* no CUDA device / CUDA error (see rqb_last_error) */
#define RQB_E_TOOBIG (-101)  /* program needs more rows than a row reference can address */

#define RQB_NO_ROW 0xFFFFFFFFu

const char *rqb_last_error(void);
int rqb_device_count(void);
/* selects the device new solvers/matrices are created on (process-wide: CUDA's
 * current device is per thread, the library binds calling threads itself) */
int rqb_set_device(int dev);
unsigned long long rqb_kernel_launches(void);
/* bytes copied host->device / device->host by this library so far */
void rqb_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h);

/* slow-path events so far: {fresh pinned allocations, fresh device allocations, arena
 * regrowths, solver contexts created}.  Each takes the driver's global lock for
 * milliseconds; in steady state the counters must stay flat. */
void rqb_slow_path_counters(unsigned long long out[4]);

/* host-side time accounting of the nanorq.h layer (enabled by the environment
 * variable NANORQ_B200_PROFILE=1; off by default).  seconds[k] = time summed over
 * all threads in slot k, names from rqb_host_profile_name(k) (NULL past the end). */
void rqb_host_profile(double *seconds, int n);
const char *rqb_host_profile_name(int k);
void rqb_host_profile_reset(void);

/* ---- RFC 6330 construction helpers (host, integer only) */
typedef struct {
  int Kprime, S, H, W, L, P, P1, U, B, J;
} rqb_block_params;
int rqb_block_params_init(int K, rqb_block_params *out);              /* lib/params.c:21 */
int rqb_lt_row_indices(int K, uint32_t isi, uint32_t *out /*>=40*/); /* lib/params.c:47 */

/* ---- one source block on the device ------------------------------------ */
typedef struct rqb_solver rqb_solver;

/* K source symbols of T bytes; room for max_in input rows (>= K) and max_out
 * emitted symbols per solve. */
int rqb_solver_create(rqb_solver **out, int K, size_t T, uint32_t max_in, uint32_t max_out);
/* K_params selects K' (the reference uses block 0's parameters for every block of
 * an object, lib/nanorq.c:289,372, so a shorter block may be padded further) */
int rqb_solver_create_ex(rqb_solver **out, int K, int K_params, size_t T, uint32_t max_in, uint32_t max_out);
/* same, on CUDA device `dev` (independent source blocks shard over the GPUs of a box with no
 * exchange between them; dev < 0 = the process-wide default of rqb_set_device) */
int rqb_solver_create_on(rqb_solver **out, int dev, int K, int K_params, size_t T, uint32_t max_in,
                         uint32_t max_out);
int rqb_solver_device(const rqb_solver *s);
/* which flavour of the solve program the planner is asked for (rqb_program.h): AUTO = the
 * shared-memory flavour when the block's rows fit a CTA's shared memory (lowest latency for a
 * block on its own: DRAM traffic is the compulsory traffic), HBM = always the HBM flavour
 * (highest throughput when many blocks are launched together) */
#define RQB_FLAVOUR_AUTO 0
#define RQB_FLAVOUR_HBM 1
void rqb_solver_set_flavour(rqb_solver *s, int flavour);
/* copy n rows from CALLER memory (rows src_pitch apart, T bytes each) into input rows
 * [first, first+n), asynchronously; plain DMA when the memory is page-locked
 * (rqb_host_alloc / rqb_host_register), staged by the driver otherwise.  The caller's rows
 * must stay valid until the next rqb_solver_sync (or any call that waits). */
int rqb_solver_upload_rows(rqb_solver *s, uint32_t first, uint32_t n, const uint8_t *src, size_t src_pitch);
/* copy rows [first, first+n) of a row space (0 = input, 1 = emitted symbols, 2 = intermediate
 * symbols) into caller memory; wait != 0 waits for the copy */
int rqb_solver_fetch_rows(rqb_solver *s, int space, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch,
                          int wait);
/* emitted-symbol row sym_row[k] = input row in_row[k], k < n, on the device (a decoder places the
 * source symbols it received into the block image next to the recovered ones) */
int rqb_solver_copy_in_to_sym(rqb_solver *s, const uint32_t *sym_row, const uint32_t *in_row, uint32_t n);
/* where the u x u Schur system of a block is eliminated while its program is built (the reference's
 * precode_matrix_solve_gf2 / _solve_gf256, lib/precode.c:264-315): 0 = on the device
 * (rqb_usolve_kernel: one CTA, warp-level min reductions for the pivot search) when the system has at
 * least 256 columns -- smaller ones are done faster on the host than a kernel round trip takes --,
 * 1 = always on the host, 2 = on the device whenever it fits.  Same pivots and results either way. */
void rqb_set_usolve_mode(int mode);
/* page-locked host memory for symbol buffers: copies to and from it are DMA without a staging copy */
void *rqb_host_alloc(size_t bytes);
void rqb_host_release(void *p);
int rqb_host_pin(void *p, size_t bytes);   /* page-lock memory the caller allocated */
int rqb_host_unpin(void *p);
/* destroy never blocks: the context (stream, events, pinned and device buffers) is parked
 * as it is -- work may still be queued on its stream -- for the next create of the same
 * shape, because CUDA object creation is too slow for blocks that come and go at wire rate;
 * whoever takes a parked context waits for its stream first.  Parked contexts and pooled
 * buffers count against a byte limit (default 48 GiB, NANORQ_B200_CACHE_MB or
 * rqb_set_cache_limit); beyond it the oldest parked contexts are handed back to the driver
 * (cudaFree / cudaFreeHost).  The cache of encoder programs is limited to 48 entries (LRU).
 * rqb_release_cached() hands EVERYTHING cached back: parked contexts, pooled buffers, cached
 * encoder programs (host and device copies), recycled plan objects and per-K' matrices; it
 * must not run concurrently with other calls into the library. */
void rqb_solver_destroy(rqb_solver *s);
void rqb_release_cached(void);
void rqb_set_cache_limit(size_t bytes);
void rqb_cache_stats(size_t *cached_bytes, size_t *limit_bytes);
/* free / total memory of the library's default device (cudaMemGetInfo) */
int rqb_device_mem_info(size_t *free_bytes, size_t *total_bytes);
/* pinned host staging area for the input rows: max_in rows of rqb_solver_pitch bytes */
uint8_t *rqb_solver_staging(rqb_solver *s);
size_t rqb_solver_pitch(const rqb_solver *s);
/* copy staging rows [first, first+n) to the device (asynchronous) */
int rqb_solver_upload(rqb_solver *s, uint32_t first, uint32_t n);

typedef struct {
  int overhead;            /* LT rows beyond K'                                          */
  const uint32_t *isi;     /* [K'+overhead] internal symbol id of LT row k               */
  const uint32_t *in_row;  /* [K'+overhead] staging row holding it, RQB_NO_ROW = zeros   */
  int want_c;              /* keep the L intermediate symbols on the device              */
  uint32_t n_out;          /* symbols to emit with the solve                             */
  const uint32_t *out_isi; /* [n_out]                                                    */
  const uint32_t *out_row; /* optional [n_out]: row of the emitted-symbol space symbol k */
                           /* goes to (default k); lets a decoder recover straight into  */
                           /* the block image it hands back with one copy                */
} rqb_solve_request;

/* analyse the block and stage the program on the device.  Returns RQB_NEED_MORE
 * when the constraint matrix is singular.  encode_plan != 0 selects (and caches
 * process-wide, per K) the plan of an encoder: isi = identity, rows 0..K-1. */
int rqb_solver_plan(rqb_solver *s, const rqb_solve_request *req);
int rqb_solver_plan_encode(rqb_solver *s, int want_c, uint32_t n_repair_with_solve);
/* rqb_solver_plan for n independent blocks on up to nthreads host threads (the analysis of one block
 * is single-threaded: ~2 ms at K=4096, ~40 ms at K=56403; blocks are independent, so a caller that
 * holds several plans them side by side).  rc[k] receives what rqb_solver_plan(solvers[k], &reqs[k])
 * returns; the function returns the number of blocks whose rc is 0.  nthreads <= 1: in the calling
 * thread. */
int rqb_solver_plan_batch(rqb_solver **solvers, const rqb_solve_request *reqs, int n, int nthreads, int *rc);
/* host threads nanorq_repair_blocks uses for the analysis of its blocks (default 1) */
void rqb_set_plan_threads(int n);
int rqb_get_plan_threads(void);
/* launch the solve kernel (asynchronous on the solver's stream) */
int rqb_solver_run(rqb_solver *s);
/* LT-combine further symbols from the intermediate symbols kept by want_c */
int rqb_solver_emit(rqb_solver *s, const uint32_t *isi, uint32_t n);
/* wait for everything queued on this solver */
int rqb_solver_sync(rqb_solver *s);
/* copy results back: emitted symbols [first, first+n) of the last run/emit, or
 * intermediate symbols; dst rows are dst_pitch apart, T bytes each */
int rqb_solver_fetch_syms(rqb_solver *s, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch);
/* queue the copy of emitted symbols into the pinned mirror without waiting;
 * rqb_solver_sync() completes it */
int rqb_solver_fetch_syms_async(rqb_solver *s, uint32_t first, uint32_t n);
int rqb_solver_fetch_c(rqb_solver *s, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch);
/* pinned host mirror of the emitted symbols (valid after fetch with dst == NULL) */
const uint8_t *rqb_solver_sym_mirror(rqb_solver *s);
/* device time of the last rqb_solver_run in milliseconds (CUDA events); needs
 * rqb_solver_set_timing(s, 1) before the run (off by default: two driver calls per block) */
void rqb_solver_set_timing(rqb_solver *s, int on);
int rqb_solver_last_kernel_ms(rqb_solver *s, float *ms);

typedef struct {
  int i, u, nb, rho, nfree, levels_fwd, n_levels, n_tasks, n_pages;
  size_t n_srcs, n_gf_srcs, n_horner, nnz; /* XOR sources, GF(256) sources, scan entries, matrix non-zeros */
  double t_matrix, t_peel, t_dense, t_emit; /* host seconds per planning phase */
  uint32_t n_ws_rows; /* working rows the program uses in HBM */
  int n_parts;        /* partial sums scheduled off the critical path */
  int slice_bytes;    /* column slice one CTA owns: the HBM flavour's default (see rqb_batch_slice_bytes) or the
                         shared-memory flavour's slot width */
  int smem;           /* 1: shared-memory flavour of the program (rows live in shared-memory slots) */
  uint32_t n_slots, tab_bits; /* shared-memory flavour: slots per CTA, inactive symbols per XOR-table group */
} rqb_solver_stats;
int rqb_solver_get_stats(const rqb_solver *s, rqb_solver_stats *out);

/* run several solvers' pending programs as ONE kernel launch (gridDim.y = n);
 * all must share T and device.  Asynchronous on solvers[0]'s stream.  Ordering across the
 * members' own streams is kept on the device with events: the launch waits for whatever a
 * member had queued (uploads), and anything a member queues afterwards (fetch, emit, upload,
 * a later run) waits for the launch -- callers need no extra synchronisation. */
int rqb_solver_run_batch(rqb_solver **solvers, int n);
/* bytes of every symbol one CTA owns (64, 128 or 256) in a launch over nblocks blocks of
 * T-byte symbols: wide slices for big batches, narrow ones for a block on its own */
int rqb_batch_slice_bytes(int nblocks, size_t T);
/* same, on the stream of `owner` (so that several batches queue behind each other) */
int rqb_solver_run_batch_on(rqb_solver **solvers, int n, rqb_solver *owner);
/* CUDA-event timing of whatever is queued on this solver's stream between
 * mark(s,0) and mark(s,1) */
int rqb_solver_mark(rqb_solver *s, int end);
int rqb_solver_marked_ms(rqb_solver *s, float *ms);

/* host-only: build the plan and hand back the raw program (tests / tooling);
 * free with rqb_plan_blob_free. */
typedef struct {
  uint32_t n_ws_rows, n_pages, page_bytes;
  /* arena layout the program addresses: first row of the spaces IN, SYM, C, WS; the
   * all-zero row; total rows.  IN is sized to the largest in_row of the request + 1,
   * SYM to n_out. */
  uint32_t row0[4], zero_row, n_rows;
  const uint8_t *pages;
  rqb_solver_stats stats;
  void *opaque;
  /* shared-memory flavour of the program (rqb_program.h): rows live in n_slots slots of
   * slice_bytes per CTA, XOR tables over tab_bits inactive symbols; smem == 0: HBM flavour */
  int smem;
  uint32_t slice_bytes, n_slots, tab_bits;
} rqb_plan_blob;
int rqb_plan_blob_build(int K_params, const rqb_solve_request *req, rqb_plan_blob *out);
/* smem_budget: bytes of shared memory a CTA may use for row slots (0 = HBM flavour only;
 * rqb_smem_budget() = what the solve kernel has on this build) */
int rqb_plan_blob_build_ex(int K_params, const rqb_solve_request *req, uint32_t smem_budget, rqb_plan_blob *out);
uint32_t rqb_smem_budget(void);
void rqb_plan_blob_free(rqb_plan_blob *b);

/* ---- batched row operations out of HBM --------------------------------- */
/* binary-compatible with the reference's sched_op (include/sched.h:6-10) */
typedef struct {
  uint8_t beta; /* >=1: D[i] ^= beta*D[j] (oaxpy/oaddrow) ; 0: D[i] *= (uint8_t)j (oscal) */
  uint32_t i;
  uint32_t j;
} rqb_op;

typedef struct rqb_matrix rqb_matrix; /* rows x T bytes resident in HBM */
int rqb_matrix_create(rqb_matrix **out, size_t rows, size_t T);
void rqb_matrix_destroy(rqb_matrix *m);
size_t rqb_matrix_pitch(const rqb_matrix *m);
int rqb_matrix_upload(rqb_matrix *m, size_t first, size_t n, const uint8_t *src, size_t src_pitch);
int rqb_matrix_download(rqb_matrix *m, size_t first, size_t n, uint8_t *dst, size_t dst_pitch);
int rqb_matrix_fill_random(rqb_matrix *m, uint64_t seed); /* device-side fill, for benchmarks */
/* one batch of mutually independent ops = one kernel launch */
int rqb_rowops_apply(rqb_matrix *m, const rqb_op *ops, size_t n);
/* same, ops already on the device (rqb_ops_upload), with CUDA-event timing */
typedef struct rqb_oplist rqb_oplist;
int rqb_ops_upload(rqb_oplist **out, const rqb_op *ops, size_t n);
void rqb_ops_free(rqb_oplist *l);
int rqb_rowops_apply_dev(rqb_matrix *m, const rqb_oplist *l, int repeats, float *ms_total);

/* replay a reference-format schedule: ops applied in the order of
 * precode_matrix_apply_sched (lib/precode.c:23-32) followed by the two row
 * permutations of precode_matrix_intermediate (lib/precode.c:379-389).
 * The host only orders the op list by its dependencies; every row operation runs on
 * the device.  rqb_schedule_replay turns the schedule into ONE launch of the solve
 * kernel (dependency levels inside the kernel, accumulations into one destination merged
 * into gathers); rqb_schedule_replay_stepwise is the literal form, one launch of the
 * batched row-op kernel per dependency level.  Same bytes either way. */
int rqb_schedule_replay(rqb_matrix *m, const rqb_op *ops, size_t nops, long mark0, long mark1,
                        const int *di, size_t rows, const int *c, size_t cols, float *ms_device);
/* host only: the program rqb_schedule_replay runs, over an arena [nrows matrix rows,
 * updated in place | nrows gathered rows | ZERO row]; free with rqb_plan_blob_free */
int rqb_schedule_plan_blob(size_t nrows, const rqb_op *ops, size_t nops, long mark0, long mark1, const int *di,
                           size_t rows, const int *c, size_t cols, rqb_plan_blob *out);
int rqb_schedule_replay_stepwise(rqb_matrix *m, const rqb_op *ops, size_t nops, long mark0, long mark1,
                                 const int *di, size_t rows, const int *c, size_t cols, float *ms_device);

#ifdef __cplusplus
}
#endif
#endif
