/* ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin entry points around the UNMODIFIED reference (compiled from the sources
 * where they lie under /root/reference into oracle/_ref/libnanorq_ref.so by
 * oracle/Makefile).  It only calls reference functions; no algorithm lives here.
 * Used by tests/ to pin oracle/rq_oracle.c and by bench.py's cpu_baseline /
 * --impl reference arm (kind "reference").
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nanorq.h"
#include "precode.h"
#include "tuple.h"

void decode_row(params *P, octmat *D, uint32_t row, uint8_t *ptr, size_t len);

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void ref_params(int K, int out[10]) {
  params P = params_init((uint16_t)K);
  out[0] = P.Kprime; out[1] = P.S; out[2] = P.H; out[3] = P.W; out[4] = P.L;
  out[5] = P.P; out[6] = P.P1; out[7] = P.U; out[8] = P.B; out[9] = P.J;
}

int ref_lt_indices(int K, uint32_t X, uint32_t *out) {
  params P = params_init((uint16_t)K);
  uint_vec v;
  kv_init(v);
  params_set_idxs(X, &P, &v);
  int n = (int)kv_size(v);
  for (int k = 0; k < n; k++) out[k] = kv_A(v, k);
  kv_destroy(v);
  return n;
}

uint32_t ref_rand(uint32_t y, uint32_t i, uint32_t m) { return rnd_get(y, (uint8_t)i, m); }

/* info[0]=nops info[1]=marks0 info[2]=marks1 info[3]=i info[4]=u.
 * ops_out (optional, 3 uint32 per op: beta,i,j) must hold ops_cap entries. */
static int run_solve(params *P, int overhead, const uint32_t *isi, octmat *D,
                     long info[5], uint32_t *ops_out, size_t ops_cap) {
  spmat *A = precode_matrix_gen(P, overhead);
  int nlt = P->Kprime + overhead;
  for (int k = 0; k < nlt; k++) {
    if (k < P->Kprime && isi[k] == (uint32_t)k) continue;
    int row = P->S + P->H + k;
    spmat_clear_row(A, row);
    params_set_idxs(isi[k], P, &A->idxs[row]);
  }
  schedule *S = precode_matrix_invert(P, A);
  if (!S) return 1;
  if (info) {
    info[0] = (long)kv_size(S->ops);
    info[1] = (long)S->marks[0];
    info[2] = (long)S->marks[1];
    info[3] = (long)S->i;
    info[4] = (long)S->u;
  }
  if (ops_out) {
    size_t n = kv_size(S->ops) < ops_cap ? kv_size(S->ops) : ops_cap;
    for (size_t k = 0; k < n; k++) {
      sched_op op = kv_A(S->ops, k);
      ops_out[3 * k] = op.beta;
      ops_out[3 * k + 1] = op.i;
      ops_out[3 * k + 2] = op.j;
    }
  }
  precode_matrix_intermediate(P, D, S);
  sched_free(S);
  return 0;
}

/* D_in: rows x T bytes (tight), rows = L+overhead, laid out like the
 * reference's D (rows S+H.. hold symbols).  C_out: L x T tight. */
int ref_solve(int K, size_t T, int overhead, const uint32_t *isi,
              const uint8_t *D_in, uint8_t *C_out, long info[5],
              uint32_t *ops_out, size_t ops_cap) {
  params P = params_init((uint16_t)K);
  octmat D = OM_INITIAL;
  size_t rows = (size_t)P.L + (size_t)overhead;
  om_resize(&D, rows, T);
  for (size_t r = 0; r < rows; r++) memcpy(om_R(D, r), D_in + r * T, T);
  int rc = run_solve(&P, overhead, isi, &D, info, ops_out, ops_cap);
  if (rc == 0)
    for (int r = 0; r < P.L; r++) memcpy(C_out + (size_t)r * T, om_R(D, r), T);
  om_destroy(&D);
  return rc;
}

void ref_lt_row(int K, size_t T, const uint8_t *C, uint32_t isi, uint8_t *out) {
  params P = params_init((uint16_t)K);
  octmat D = OM_INITIAL;
  om_resize(&D, P.L, T);
  for (int r = 0; r < P.L; r++) memcpy(om_R(D, r), C + (size_t)r * T, T);
  decode_row(&P, &D, isi, out, T);
  om_destroy(&D);
}

/* Public-API block round trip pieces with wall-clock timing (seconds), the
 * scope benchmark.c times: generate_symbols (benchmark.c:101-109) and
 * repair_block (benchmark.c:143-151).  esis: the ESIs to emit/feed, in order. */
int ref_encode_api(size_t K, size_t T, const uint8_t *payload, const uint32_t *esis,
                   size_t n, uint8_t *syms_out, uint64_t oti[2], double *t_gen,
                   double *t_emit, int precalc) {
  size_t F = K * T;
  nanorq *rq = nanorq_encoder_new_ex(F, (uint16_t)T, (uint16_t)K, 0, 8);
  if (!rq) return -1;
  struct ioctx *io = ioctx_from_mem(payload, F);
  oti[0] = nanorq_oti_common(rq);
  oti[1] = nanorq_oti_scheme_specific(rq);
  if (precalc) nanorq_precalculate(rq);
  double t0 = now_s();
  bool ok = nanorq_generate_symbols(rq, 0, io);
  double t1 = now_s();
  int rc = ok ? 0 : 1;
  for (size_t k = 0; k < n && rc == 0; k++)
    if (nanorq_encode(rq, syms_out + k * T, esis[k], 0, io) != T) rc = 2;
  double t2 = now_s();
  if (t_gen) *t_gen = t1 - t0;
  if (t_emit) *t_emit = t2 - t1;
  io->destroy(io);
  nanorq_free(rq);
  return rc;
}

int ref_decode_api(const uint64_t oti[2], size_t T, const uint32_t *esis,
                   const uint8_t *syms, size_t n, uint8_t *out, size_t out_len,
                   double *t_add, double *t_repair) {
  nanorq *rq = nanorq_decoder_new(oti[0], (uint32_t)oti[1]);
  if (!rq) return -1;
  struct ioctx *io = ioctx_from_mem(out, out_len);
  double t0 = now_s();
  for (size_t k = 0; k < n; k++)
    if (nanorq_decoder_add_symbol(rq, (void *)(syms + k * T), nanorq_tag(0, esis[k]), io) ==
        NANORQ_SYM_ERR) {
      io->destroy(io);
      nanorq_free(rq);
      return -2;
    }
  double t1 = now_s();
  bool ok = nanorq_repair_block(rq, io, 0);
  double t2 = now_s();
  if (t_add) *t_add = t1 - t0;
  if (t_repair) *t_repair = t2 - t1;
  io->destroy(io);
  nanorq_free(rq);
  return ok ? 0 : 1;
}
