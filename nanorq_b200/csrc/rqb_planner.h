/* rqb_planner.h -- host-side construction of the solve program.
 *
 * Replaces, for the B200 build, the reference's precode_matrix_gen /
 * precode_matrix_invert (lib/precode.c:90-97,347-377) and the permutation half
 * of precode_matrix_intermediate (lib/precode.c:379-389): it analyses the
 * sparse constraint matrix of one source block and emits the device program
 * described in rqb_program.h.  No symbol data is touched on the host.
 */
#ifndef RQB_PLANNER_H
#define RQB_PLANNER_H

#include <stddef.h>
#include <stdint.h>

#include "rqb_rfc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int K;                   /* source symbols of the block (selects K')                      */
  int overhead;            /* LT rows beyond K' (received repair symbols - missing symbols) */
  const uint32_t *isi;     /* [K'+overhead] internal symbol id of the LT row k              */
  const uint32_t *in_row;  /* [K'+overhead] input row (RQB_SP_IN) with that symbol's bytes, */
                           /*               RQB_ROW_NONE for an all-zero symbol (padding)   */
  int want_c;              /* emit all L intermediate symbols to c_out (row = index)        */
  int n_out;               /* encoding symbols to emit to sym_out (row k = out_isi[k])      */
  const uint32_t *out_isi; /* [n_out] internal symbol ids                                   */
  uint32_t in_rows;        /* rows the arena reserves for the input space  (> every in_row) */
  uint32_t sym_rows;       /* rows the arena reserves for emitted symbols  (>= n_out)       */
  uint8_t *pages_buf;      /* optional: write the program here (e.g. pinned memory the DMA    */
  size_t pages_buf_cap;    /* engine reads) instead of the plan's own buffer; if it is too     */
                           /* small the plan's buffer is used (plan->pages tells which)        */
  const uint32_t *out_row; /* optional [n_out]: row of the SYM space symbol k is written to (default k) */
  uint32_t smem_budget;    /* bytes of shared memory a CTA may use for row slots: > 0 asks for */
                           /* the shared-memory flavour of the program (rqb_program.h) when    */
                           /* the block's live rows fit, 0 for the HBM flavour                 */
} rqb_plan_request;

typedef struct {
  int i, u;            /* peeled rows / inactivated columns (cf. schedule.i, .u)   */
  int nb, rho, nfree;  /* binary residual rows, their GF(2) rank, columns left for HDPC */
  int levels_fwd;      /* dependency depth of the sparse triangular solve          */
  int n_parts;         /* partial sums that keep long rows off the critical path   */
  int n_levels, n_tasks, n_pages;
  size_t n_srcs, n_gf_srcs, n_horner;
  size_t nnz;
  double t_matrix, t_peel, t_dense, t_emit; /* seconds */
} rqb_plan_stats;

typedef struct rqb_plan {
  rqb_params P;
  int K, overhead;
  uint32_t n_ws_rows;  /* working rows (RQB_SP_WS) the program uses             */
  uint32_t row0[4];    /* first arena row of each space (rqb_program.h)         */
  uint32_t n_rows;     /* arena rows the program addresses: row0[WS] + n_ws_rows */
  uint32_t zero_row;   /* the all-zero row (the solver clears it, nothing writes it) */
  uint32_t n_pages;
  uint8_t *pages;      /* n_pages * RQB_PAGE_BYTES: own_pages or the request's pages_buf */
  uint8_t *own_pages;  /* the plan's own buffer (plans are recycled)            */
  size_t pages_cap;    /* bytes allocated behind own_pages                      */
  struct rqb_plan *next_free;
  uint32_t n_c_rows;   /* rows written to c_out (L or 0)                        */
  int smem;            /* 1: shared-memory flavour (rows live in slots of slice_bytes)   */
  uint32_t slice_bytes; /* smem flavour: column slice the slot budget was computed for   */
  uint32_t n_slots;    /* smem flavour: slots a CTA needs (slot 0 = zeros)               */
  uint32_t tab_bits;   /* smem flavour: inactive symbols per table group                 */
  uint32_t n_out;
  rqb_plan_stats st;
} rqb_plan;

/* 0 = ok, 1 = matrix rank < L (more symbols needed; the reference's
 * precode_matrix_invert returns NULL, lib/precode.c:368-370), <0 = bad request */
int rqb_plan_build(const rqb_plan_request *req, rqb_plan **out);
void rqb_plan_free(rqb_plan *p);
void rqb_plan_pool_drain(void); /* really free recycled plans and cached per-K' matrices */

/* one device program from an ordered sequence of reference-format row operations
 * (rqb_rowop[nops]) followed by a row gather; see rqb_planner.c */
int rqb_plan_from_schedule(const void *ops, size_t nops, uint32_t nrows, const uint32_t *gather_map, uint32_t base,
                           uint32_t out_base, uint32_t zero_row, rqb_plan **out);

/* The elimination of the u x u Schur system (step 3d/3e of rqb_plan_build: GF(2) Gauss-Jordan on the
 * binary residual rows with a tracked transformation, then the H x nfree GF(256) system of the HDPC
 * rows; the reference's precode_matrix_solve_gf2 / _solve_gf256, lib/precode.c:264-315) can run on
 * the device: rqb_solver.c installs this hook when a GPU is present and rqb_plan_build calls it for
 * blocks whose system is large enough to pay for the round trip (rqb_set_usolve_mode).  The kernel
 * picks the same pivots as the host code (lowest unused row with the bit set), so programs are
 * identical either way.  Returns 0 = solved, 1 = singular (rank < L), -1 = not available (host code
 * runs instead). */
typedef struct {
  int nb, U, uw, nbw, H;
  size_t sh_stride;    /* bytes between the HDPC Schur rows in Sh */
  uint64_t *Sb;        /* [nb x uw]  in: binary Schur rows; out: reduced rows                  */
  uint64_t *Tb;        /* [nb x nbw] out: the transformation (in: ignored, starts as identity)  */
  const uint8_t *Sh;   /* [H x sh_stride] HDPC Schur rows (bytes)                               */
  int *pivrow;         /* [U]  out: pivot row of column t or -1                                 */
  uint8_t *TQ;         /* [H x H] out                                                          */
  int *qrow_of_f;      /* [<= H] out                                                           */
  int nfree, rho;      /* out */
} rqb_usolve_io;
extern int (*rqb_plan_usolve_hook)(rqb_usolve_io *io);
/* 0 = auto (device when U >= 256 and a hook is installed), 1 = always host, 2 = device whenever possible */
extern int rqb_plan_usolve_mode;

int rqb_params_init(int K, rqb_params *P);
/* host-side Tuple / index helpers over the built-in tables */
int rqb_host_lt_indices(const rqb_params *P, uint32_t X, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
