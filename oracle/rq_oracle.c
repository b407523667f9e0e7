/* rq_oracle.c -- TEST INFRASTRUCTURE ONLY (see rq_oracle.h).
 *
 * Scalar C99 restatement of the reference hot path.  Every function cites the
 * reference file:line whose behaviour it restates.  Data structures are this
 * project's own (flat CSR arrays, flat bit/byte matrices); the *algorithm* --
 * row order, LIFO peeling buckets, pivot order, recorded op order -- is the
 * reference's, so op lists match the compiled reference 1:1.
 */
#include "rq_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "rfc6330_tables.h"

/* ------------------------------------------------------------------ GF(256) */
/* deps/oblas/tablegen.c:10,31-52 : poly 0x11D, generator 2, LOG[0]=255,
 * EXP has 510 live entries so LOG[a]+LOG[b] needs no reduction (oblas.h:9). */
static uint8_t gf_exp_t[512], gf_log_t[256];
static int gf_ready;

static void gf_init(void) {
  if (gf_ready) return;
  unsigned x = 1;
  for (int e = 0; e < 255; e++) {
    gf_exp_t[e] = (uint8_t)x;
    gf_log_t[x] = (uint8_t)e;
    x <<= 1;
    if (x & 0x100) x ^= 0x11D;
  }
  for (int e = 255; e < 512; e++) gf_exp_t[e] = gf_exp_t[e - 255];
  gf_log_t[0] = 255;
  gf_ready = 1;
}

uint8_t orc_gf_exp(int e) {
  gf_init();
  return gf_exp_t[e % 255];
}

uint8_t orc_gf_mul(uint8_t a, uint8_t b) {
  gf_init();
  if (!a || !b) return 0;
  return gf_exp_t[gf_log_t[a] + gf_log_t[b]];
}

uint8_t orc_gf_inv(uint8_t a) {
  gf_init();
  if (!a) return 0;
  return gf_exp_t[255 - gf_log_t[a]];
}

/* oaxpy: deps/oblas/oblas_classic.c (u==0 no-op, u==1 plain xor) and
 * oblas_avx.c:43-74 */
void orc_row_axpy(uint8_t *dst, const uint8_t *src, size_t n, uint8_t u) {
  gf_init();
  if (u == 0) return;
  if (u == 1) {
    for (size_t k = 0; k < n; k++) dst[k] ^= src[k];
    return;
  }
  unsigned lu = gf_log_t[u];
  for (size_t k = 0; k < n; k++)
    if (src[k]) dst[k] ^= gf_exp_t[gf_log_t[src[k]] + lu];
}

/* oscal: oblas_avx.c:91-114 -- u<2 is a no-op (u==0 does NOT zero the row) */
void orc_row_scal(uint8_t *dst, size_t n, uint8_t u) {
  gf_init();
  if (u < 2) return;
  unsigned lu = gf_log_t[u];
  for (size_t k = 0; k < n; k++)
    dst[k] = dst[k] ? gf_exp_t[gf_log_t[dst[k]] + lu] : 0;
}

/* precode_matrix_apply_op: lib/precode.c:15-21 */
static void apply_one(uint8_t *D, size_t pitch, size_t T, orc_op op) {
  if (op.beta)
    orc_row_axpy(D + (size_t)op.i * pitch, D + (size_t)op.j * pitch, T, op.beta);
  else
    orc_row_scal(D + (size_t)op.i * pitch, T, (uint8_t)op.j);
}

void orc_apply_ops(uint8_t *D, size_t pitch, size_t T, const orc_op *ops,
                   size_t nops) {
  for (size_t k = 0; k < nops; k++) apply_one(D, pitch, T, ops[k]);
}

/* --------------------------------------------------- params / rand / tuple */
static int is_prime(int n) { /* lib/params.c:5-19 */
  if (n < 2) return 0;
  for (int q = 2; q * q <= n; q++)
    if (n % q == 0) return 0;
  return 1;
}

int orc_params_init(int K, orc_params *o) { /* lib/params.c:21-45 */
  if (K < 1 || K > RQB_MAX_K) return -1;
  int idx = 0;
  while (rqb_kprime_table[idx].kprime < K) idx++;
  const rqb_kprime_row *r = &rqb_kprime_table[idx];
  o->Kprime = r->kprime;
  o->J = r->j;
  o->S = r->s;
  o->H = r->h;
  o->W = r->w;
  o->L = o->Kprime + o->S + o->H;
  o->P = o->L - o->W;
  o->U = o->P - o->H;
  o->B = o->W - o->S;
  o->P1 = o->P;
  while (!is_prime(o->P1)) o->P1++;
  return 0;
}

uint32_t orc_rand(uint32_t y, uint32_t i, uint32_t m) { /* lib/rand.c:183-190 */
  uint32_t v = rqb_rand_v[0][(y + i) & 0xff] ^
               rqb_rand_v[1][((y >> 8) + i) & 0xff] ^
               rqb_rand_v[2][((y >> 16) + i) & 0xff] ^
               rqb_rand_v[3][((y >> 24) + i) & 0xff];
  return v % m;
}

static uint32_t degree_of(uint32_t v, int W) { /* lib/tuple.c:13-19 */
  for (int d = 0; d < 31; d++)
    if (v < rqb_degree_cdf[d]) return (uint32_t)(d < W - 2 ? d : W - 2);
  return 0;
}

orc_tuple orc_tuple_gen(const orc_params *P, uint32_t X) { /* lib/tuple.c:21-43 */
  orc_tuple t;
  uint64_t A = 53591 + (uint64_t)P->J * 997;
  if ((A & 1) == 0) A++;
  uint64_t B = 10267 * ((uint64_t)P->J + 1);
  uint32_t y = (uint32_t)(B + (uint64_t)X * A);
  uint32_t v = orc_rand(y, 0, 1u << 20);
  t.d = degree_of(v, P->W);
  t.a = 1 + orc_rand(y, 1, (uint32_t)P->W - 1);
  t.b = orc_rand(y, 2, (uint32_t)P->W);
  t.d1 = (t.d < 4) ? 2 + orc_rand(X, 3, 2) : 2;
  t.a1 = 1 + orc_rand(X, 4, (uint32_t)P->P1 - 1);
  t.b1 = orc_rand(X, 5, (uint32_t)P->P1);
  return t;
}

int orc_lt_indices(const orc_params *P, uint32_t X, uint32_t *out) {
  /* lib/params.c:47-65 */
  orc_tuple t = orc_tuple_gen(P, X);
  int n = 0;
  uint32_t W = (uint32_t)P->W, Pp = (uint32_t)P->P, P1 = (uint32_t)P->P1;
  out[n++] = t.b;
  for (uint32_t j = 1; j < t.d; j++) {
    t.b = (t.b + t.a) % W;
    out[n++] = t.b;
  }
  while (t.b1 >= Pp) t.b1 = (t.b1 + t.a1) % P1;
  out[n++] = W + t.b1;
  for (uint32_t j = 1; j < t.d1; j++) {
    t.b1 = (t.b1 + t.a1) % P1;
    while (t.b1 >= Pp) t.b1 = (t.b1 + t.a1) % P1;
    out[n++] = W + t.b1;
  }
  return n;
}

void orc_lt_row(const orc_params *P, const uint8_t *C, size_t pitch,
                uint32_t isi, uint8_t *out, size_t T) {
  /* decode_row: lib/nanorq.c:184-204 */
  uint32_t idx[40];
  int n = orc_lt_indices(P, isi, idx);
  memset(out, 0, T);
  for (int k = 0; k < n; k++) {
    const uint8_t *row = C + (size_t)idx[k] * pitch;
    for (size_t b = 0; b < T; b++) out[b] ^= row[b];
  }
}

/* ------------------------------------------------------ sparse 0/1 matrix */
typedef struct {
  int rows, cols;
  int *ptr; /* rows+1 */
  int *idx; /* nnz, in the reference's push order */
} csr;

static void csr_free(csr *m) {
  if (!m) return;
  free(m->ptr);
  free(m->idx);
  free(m);
}

/* precode_matrix_gen: lib/precode.c:34-58,85-97 (+ patch_precode_matrix,
 * lib/nanorq.c:527-547, expressed through isi[]).  HDPC rows S..S+H-1 are left
 * empty exactly like the reference. */
static csr *matrix_gen(const orc_params *P, int overhead, const uint32_t *isi) {
  int rows = P->L + overhead, S = P->S, H = P->H, B = P->B, W = P->W;
  csr *A = calloc(1, sizeof(*A));
  A->rows = rows;
  A->cols = P->L;
  A->ptr = calloc((size_t)rows + 1, sizeof(int));
  int *cnt = calloc((size_t)rows, sizeof(int));
  uint32_t tmp[40];
  /* pass 1: counts */
  for (int col = 0; col < B; col++) {
    int sub = col / S;
    cnt[col % S]++;
    cnt[(col + sub + 1) % S]++;
    cnt[(col + 2 * (sub + 1)) % S]++;
  }
  for (int r = 0; r < S; r++) cnt[r] += 1 + 2;
  int nlt = P->Kprime + overhead;
  for (int k = 0; k < nlt; k++) cnt[S + H + k] = orc_lt_indices(P, isi[k], tmp);
  for (int r = 0; r < rows; r++) A->ptr[r + 1] = A->ptr[r] + cnt[r];
  A->idx = malloc(sizeof(int) * (size_t)(A->ptr[rows] ? A->ptr[rows] : 1));
  int *cur = malloc(sizeof(int) * (size_t)rows);
  memcpy(cur, A->ptr, sizeof(int) * (size_t)rows);
  /* pass 2: same push order as the reference */
  for (int col = 0; col < B; col++) { /* make_LDPC1 :39-49 */
    int sub = col / S;
    A->idx[cur[col % S]++] = col;
    A->idx[cur[(col + sub + 1) % S]++] = col;
    A->idx[cur[(col + 2 * (sub + 1)) % S]++] = col;
  }
  for (int r = 0; r < S; r++) A->idx[cur[r]++] = B + r; /* identity :34-37 */
  for (int r = 0; r < S; r++) {                         /* make_LDPC2 :51-58 */
    A->idx[cur[r]++] = W + r % P->P;
    A->idx[cur[r]++] = W + (r + 1) % P->P;
  }
  for (int k = 0; k < nlt; k++) { /* make_G_ENC :85-88 */
    int n = orc_lt_indices(P, isi[k], tmp);
    for (int q = 0; q < n; q++) A->idx[cur[S + H + k]++] = (int)tmp[q];
  }
  free(cnt);
  free(cur);
  return A;
}

/* spmat_transpose: lib/spmat.c:38-46 (column lists hold rows ascending) */
static csr *csr_transpose(const csr *A) {
  csr *T = calloc(1, sizeof(*T));
  T->rows = A->cols;
  T->cols = A->rows;
  T->ptr = calloc((size_t)T->rows + 1, sizeof(int));
  int nnz = A->ptr[A->rows];
  T->idx = malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
  for (int k = 0; k < nnz; k++) T->ptr[A->idx[k] + 1]++;
  for (int c = 0; c < T->rows; c++) T->ptr[c + 1] += T->ptr[c];
  int *cur = malloc(sizeof(int) * (size_t)T->rows);
  memcpy(cur, T->ptr, sizeof(int) * (size_t)T->rows);
  for (int r = 0; r < A->rows; r++)
    for (int k = A->ptr[r]; k < A->ptr[r + 1]; k++) T->idx[cur[A->idx[k]]++] = r;
  free(cur);
  return T;
}

/* ------------------------------------------------------------- schedule */
static orc_sched *sched_new(int rows, int cols) { /* lib/sched.c:3-27 */
  orc_sched *S = calloc(1, sizeof(*S));
  S->rows = rows;
  S->cols = cols;
  S->c = malloc(sizeof(int) * (size_t)cols);
  S->ci = malloc(sizeof(int) * (size_t)cols);
  S->d = malloc(sizeof(int) * (size_t)rows);
  S->di = malloc(sizeof(int) * (size_t)rows);
  for (int j = 0; j < cols; j++) S->c[j] = S->ci[j] = j;
  for (int r = 0; r < rows; r++) S->d[r] = S->di[r] = r;
  S->cap = 3 * (size_t)cols + 16;
  S->ops = malloc(sizeof(orc_op) * S->cap);
  return S;
}

void orc_sched_free(orc_sched *S) {
  if (!S) return;
  free(S->c);
  free(S->ci);
  free(S->d);
  free(S->di);
  free(S->ops);
  free(S);
}

static void push_op(orc_sched *S, uint32_t i, uint32_t j, uint8_t beta) {
  if (S->nops == S->cap) { /* lib/sched.c:46-49 */
    S->cap *= 2;
    S->ops = realloc(S->ops, sizeof(orc_op) * S->cap);
  }
  S->ops[S->nops].i = i;
  S->ops[S->nops].j = j;
  S->ops[S->nops].beta = beta;
  S->nops++;
}

static void swap_int(int *a, int *b) {
  int t = *a;
  *a = *b;
  *b = t;
}

/* ----------------------------------------------- hybrid working matrix U */
/* lib/wrkmat.c + deps/oblas/gf2.c : every row has a bit-packed GF(2) image;
 * rows flagged `wide` live in a byte (GF(256)) slab of 2H rows. */
typedef struct {
  int rows, cols, words, bcols;
  uint32_t *bits;
  uint8_t *bytes; /* 2H rows x bcols */
  int nbyte_rows, next_free;
  int *slot; /* -1 => GF(2) row */
  int aborted;
  int promotions;
} umat;

static umat *umat_new(int rows, int cols, int H) {
  umat *U = calloc(1, sizeof(*U));
  U->rows = rows;
  U->cols = cols;
  U->words = (cols + 31) / 32;
  if (U->words == 0) U->words = 1;
  U->bcols = U->words * 32;
  U->bits = calloc((size_t)rows * U->words, sizeof(uint32_t));
  U->nbyte_rows = 2 * H;
  U->bytes = calloc((size_t)U->nbyte_rows * U->bcols, 1);
  U->slot = malloc(sizeof(int) * (size_t)rows);
  for (int r = 0; r < rows; r++) U->slot[r] = -1;
  return U;
}

static void umat_free(umat *U) {
  if (!U) return;
  free(U->bits);
  free(U->bytes);
  free(U->slot);
  free(U);
}

static inline uint8_t umat_at(const umat *U, int r, int col) {
  /* wrkmat_at: include/wrkmat.h:21-25 */
  if (U->slot[r] >= 0) return U->bytes[(size_t)U->slot[r] * U->bcols + col];
  return (U->bits[(size_t)r * U->words + col / 32] >> (col % 32)) & 1;
}

static void umat_axpy(umat *U, int i, int j, uint8_t beta) {
  /* wrkmat_axpy: lib/wrkmat.c:76-108 */
  uint32_t *bi = U->bits + (size_t)i * U->words;
  const uint32_t *bj = U->bits + (size_t)j * U->words;
  int wi = U->slot[i] >= 0, wj = U->slot[j] >= 0;
  if (wi == wj) {
    if (wi)
      orc_row_axpy(U->bytes + (size_t)U->slot[i] * U->bcols,
                   U->bytes + (size_t)U->slot[j] * U->bcols, (size_t)U->bcols, beta);
    for (int w = 0; w < U->words; w++) bi[w] ^= bj[w];
    return;
  }
  if (wi) { /* oaxpy_b32: oblas_avx.c:149-170 */
    uint8_t *dst = U->bytes + (size_t)U->slot[i] * U->bcols;
    for (int col = 0; col < U->bcols; col++)
      if ((bj[col / 32] >> (col % 32)) & 1) dst[col] ^= beta;
    return;
  }
  /* GF(2) row receiving a GF(256) row: promote (wrkmat.c:93-105) */
  if (U->next_free >= U->nbyte_rows) {
    U->aborted = 1;
    return;
  }
  uint8_t *dst = U->bytes + (size_t)U->next_free * U->bcols;
  for (int col = 0; col < U->bcols; col++) /* gf2mat_fill: gf2.c:62-73 */
    if ((bi[col / 32] >> (col % 32)) & 1) dst[col] = 1;
  U->slot[i] = U->next_free++;
  U->promotions++;
  orc_row_axpy(dst, U->bytes + (size_t)U->slot[j] * U->bcols, (size_t)U->bcols, beta);
}

/* ------------------------------------------------------ invert: phases */
typedef struct {
  int *buf[3];
  int n[3];
} buckets;

/* precode_matrix_sort: lib/precode.c:99-109 */
static void phase_sort(const orc_params *P, const csr *A, orc_sched *S,
                       unsigned *nz) {
  int rows = A->rows;
  for (int r = 0; r < rows; r++) S->d[r] = (r + P->S + P->H) % rows;
  for (int r = 0; r < rows; r++) S->di[S->d[r]] = r;
  int lim = A->cols - P->P;
  for (int r = 0; r < rows; r++) {
    unsigned n = 0;
    for (int k = A->ptr[r]; k < A->ptr[r + 1]; k++) n += (A->idx[k] < lim);
    nz[r] = n ? n : (unsigned)A->cols;
  }
}

/* precode_row_nz_at: lib/precode.c:128-140 */
static int row_nz_at(const csr *A, int pos, int s, int e, const orc_sched *S,
                     const unsigned *nz, int at[2]) {
  int r = 0, orig = S->d[pos];
  at[0] = at[1] = e;
  for (int k = A->ptr[orig]; k < A->ptr[orig + 1] && r < (int)nz[orig]; k++) {
    int col = S->ci[A->idx[k]];
    if (col >= s && col < e) at[r++] = col;
  }
  if (at[0] > at[1]) swap_int(&at[0], &at[1]);
  return r;
}

static void nz_dec_column(const csr *AT, int col, unsigned *nz, buckets *bk) {
  /* inner loops of precode_matrix_update_nnz: lib/precode.c:156-174 */
  for (int k = AT->ptr[col]; k < AT->ptr[col + 1]; k++) {
    int row = AT->idx[k];
    unsigned v = --nz[row];
    if (v > 0 && v < 3) bk->buf[v][bk->n[v]++] = row;
  }
}

/* precode_matrix_precond: lib/precode.c:115-126,142-154,176-203 */
static void phase_precond(const orc_params *P, const csr *A, const csr *AT,
                          orc_sched *S, unsigned *nz) {
  int i = 0, u = P->P, rows = A->rows, Srows = rows - P->H, cols = A->cols;
  int *d = S->d, *di = S->di, *c = S->c, *ci = S->ci;
  buckets bk;
  size_t cap = (size_t)A->ptr[A->rows] + (size_t)rows + 8;
  for (int b = 0; b < 3; b++) {
    bk.buf[b] = malloc(sizeof(int) * cap);
    bk.n[b] = 0;
  }
  for (int r = 0; r < Srows; r++)
    if (nz[d[r]] < 3) bk.buf[nz[d[r]]][bk.n[nz[d[r]]]++] = d[r];
  while (i + u < P->L) {
    int V0 = i, Vcols = cols - i - u, chosen = Srows;
    for (int b = 1; b < 3 && chosen == Srows; b++) /* _choose :115-126 */
      while (bk.n[b] > 0) {
        int cand = bk.buf[b][--bk.n[b]];
        if (di[cand] >= V0 && nz[cand] == (unsigned)b) {
          chosen = di[cand];
          break;
        }
      }
    if (chosen >= Srows) break;
    if (V0 != chosen) {
      swap_int(&d[V0], &d[chosen]);
      swap_int(&di[d[V0]], &di[d[chosen]]);
    }
    int ones[2], Vlast = V0 + Vcols - 1; /* _swap_cols :142-154 */
    int r = row_nz_at(A, V0, V0, V0 + Vcols, S, nz, ones);
    if (ones[0] != V0) {
      swap_int(&c[V0], &c[ones[0]]);
      swap_int(&ci[c[V0]], &ci[c[ones[0]]]);
    }
    if (r == 2 && ones[1] != Vlast) {
      swap_int(&c[Vlast], &c[ones[1]]);
      swap_int(&ci[c[Vlast]], &ci[c[ones[1]]]);
    }
    nz_dec_column(AT, c[V0], nz, &bk); /* _update_nnz :156-174 */
    for (int col = 0; col < r - 1; col++)
      nz_dec_column(AT, c[V0 + Vcols - col - 1], nz, &bk);
    i++;
    u += r - 1;
  }
  for (int b = 0; b < 3; b++) free(bk.buf[b]);
  S->i = i;
  S->u = P->L - i;
}

/* precode_matrix_fwd_GE: lib/precode.c:205-219 */
static void phase_fwd(umat *U, orc_sched *S, const csr *AT, int s, int e) {
  for (int row = 0; row < S->i; row++) {
    int mv = s < row ? row : s;
    int col = S->c[row];
    for (int k = AT->ptr[col]; k < AT->ptr[col + 1]; k++) {
      int tgt = AT->idx[k], h = S->di[tgt];
      if (h > mv && h < e) {
        umat_axpy(U, tgt, S->d[row], 1);
        push_op(S, (uint32_t)tgt, (uint32_t)S->d[row], 1);
      }
    }
  }
}

/* precode_matrix_make_HDPC: lib/precode.c:60-83 ; H x (K'+S), row-major */
static uint8_t *make_hdpc(const orc_params *P) {
  int m = P->H, n = P->Kprime + P->S;
  uint8_t *M = calloc((size_t)m * n, 1);
  for (int r = 0; r < m; r++) M[(size_t)r * n + n - 1] = orc_gf_exp(r);
  for (int col = n - 2; col >= 0; col--) {
    for (int r = 0; r < m; r++)
      M[(size_t)r * n + col] = orc_gf_mul(M[(size_t)r * n + col + 1], 2);
    int b1 = (int)orc_rand((uint32_t)col + 1, 6, (uint32_t)m);
    int b2 = (b1 + (int)orc_rand((uint32_t)col + 1, 7, (uint32_t)m - 1) + 1) % m;
    M[(size_t)b1 * n + col] ^= 1;
    M[(size_t)b2 * n + col] ^= 1;
  }
  return M;
}

/* precode_matrix_fill_HDPC: lib/precode.c:232-252 */
static void phase_hdpc(const orc_params *P, umat *U, orc_sched *S) {
  int H = P->H, n = P->Kprime + P->S, u = S->u;
  uint8_t *M = make_hdpc(P);
  for (int r = 0; r < H; r++) {
    uint8_t *dst = U->bytes + (size_t)r * U->bcols;
    for (int col = 0; col < u - H; col++)
      dst[col] = M[(size_t)r * n + S->c[n - (u - H) + col]];
    dst[r + (u - H)] = 1;
    U->slot[P->S + r] = r; /* wrkmat_assign_block: lib/wrkmat.c:32-41 */
  }
  U->next_free = H;
  for (int row = 0; row < S->i; row++)
    for (int h = 0; h < H; h++) {
      uint8_t beta = M[(size_t)h * n + S->c[row]];
      if (!beta) continue;
      int tgt = S->d[U->rows - H + h];
      umat_axpy(U, tgt, S->d[row], beta);
      push_op(S, (uint32_t)tgt, (uint32_t)S->d[row], beta);
    }
  free(M);
}

/* precode_matrix_solve_gf2: lib/precode.c:264-285 */
static int phase_solve_gf2(const orc_params *P, umat *U, orc_sched *S) {
  int *d = S->d, *di = S->di, row, nzrow, rows = U->rows - P->H;
  for (row = S->i; row < P->L; row++) {
    int col = row - S->i;
    for (nzrow = row; nzrow < rows; nzrow++)
      if (umat_at(U, d[nzrow], col)) break;
    if (nzrow == rows) break;
    if (row != nzrow) {
      swap_int(&d[row], &d[nzrow]);
      swap_int(&di[d[row]], &di[d[nzrow]]);
    }
    for (int del = row + 1; del < rows; del++) {
      if (!umat_at(U, d[del], col)) continue;
      umat_axpy(U, d[del], d[row], 1);
      push_op(S, (uint32_t)d[del], (uint32_t)d[row], 1);
    }
  }
  return row;
}

/* precode_matrix_solve_gf256: lib/precode.c:287-315 */
static int phase_solve_gf256(const orc_params *P, umat *U, orc_sched *S) {
  int *d = S->d, *di = S->di, row, nzrow, rows = U->rows;
  for (row = S->i; row < P->L; row++) {
    int col = row - S->i;
    uint8_t beta = 0;
    for (nzrow = row; nzrow < rows; nzrow++) {
      beta = umat_at(U, d[nzrow], col);
      if (beta) break;
    }
    if (nzrow == rows) break;
    if (row != nzrow) {
      swap_int(&d[row], &d[nzrow]);
      swap_int(&di[d[row]], &di[d[nzrow]]);
    }
    if (beta > 1) { /* wrkmat_scal: lib/wrkmat.c:110-118 (GF(256) rows only) */
      uint8_t inv = orc_gf_inv(beta);
      if (U->slot[d[row]] < 0) {
        U->aborted = 1;
        return row;
      }
      orc_row_scal(U->bytes + (size_t)U->slot[d[row]] * U->bcols, (size_t)U->bcols, inv);
      push_op(S, (uint32_t)d[row], inv, 0);
    }
    for (int del = row + 1; del < rows; del++) {
      beta = umat_at(U, d[del], col);
      if (!beta) continue;
      umat_axpy(U, d[del], d[row], beta);
      if (U->aborted) return row;
      push_op(S, (uint32_t)d[del], (uint32_t)d[row], beta);
    }
  }
  return row;
}

/* precode_matrix_backsolve: lib/precode.c:317-334 */
static void phase_backsolve(const orc_params *P, const csr *AT, const umat *U,
                            orc_sched *S) {
  for (int row = P->L - 1; row >= S->i; row--) {
    int col = S->c[row];
    for (int k = AT->ptr[col]; k < AT->ptr[col + 1]; k++) {
      int del = S->di[AT->idx[k]];
      if (del < S->i) push_op(S, (uint32_t)S->d[del], (uint32_t)S->d[row], 1);
    }
    for (int del = S->i; del < row; del++) {
      uint8_t beta = umat_at(U, S->d[del], row - S->i);
      if (beta) push_op(S, (uint32_t)S->d[del], (uint32_t)S->d[row], beta);
    }
  }
}

orc_sched *orc_invert(const orc_params *P, int overhead, const uint32_t *isi,
                      int *status) {
  /* precode_matrix_invert: lib/precode.c:347-377 */
  gf_init();
  csr *A = matrix_gen(P, overhead, isi);
  int rows = A->rows;
  orc_sched *S = sched_new(rows, A->cols);
  unsigned *nz = malloc(sizeof(unsigned) * (size_t)rows);
  phase_sort(P, A, S, nz);
  csr *AT = csr_transpose(A);
  phase_precond(P, A, AT, S, nz);

  /* precode_matrix_make_U (+ fill_U): lib/precode.c:221-230,254-262 */
  umat *U = umat_new(rows, S->u, P->H);
  for (int r = 0; r < rows; r++)
    for (int k = A->ptr[r]; k < A->ptr[r + 1]; k++) {
      int col = S->ci[A->idx[k]];
      if (col >= S->i) {
        int b = col - S->i;
        U->bits[(size_t)r * U->words + b / 32] |= 1u << (b % 32);
      }
    }
  phase_fwd(U, S, AT, 0, S->i);
  S->marks[0] = (long)S->nops - 1;
  phase_fwd(U, S, AT, S->i - 1, rows - P->H);

  int rank = 0, st = 0;
  if (rows - P->H >= P->L) rank = phase_solve_gf2(P, U, S);
  if (rank < P->L) {
    phase_hdpc(P, U, S);
    if (!U->aborted) rank = phase_solve_gf256(P, U, S);
    if (U->aborted)
      st = 2;
    else if (rank < P->L)
      st = 1;
  }
  if (st == 0) {
    S->marks[1] = (long)S->nops - 1;
    phase_backsolve(P, AT, U, S);
    S->promotions = U->promotions;
  }
  umat_free(U);
  free(nz);
  csr_free(A);
  csr_free(AT);
  if (status) *status = st;
  if (st) {
    orc_sched_free(S);
    return NULL;
  }
  return S;
}

size_t orc_applied_ops(const orc_sched *S, size_t *n_axpy, size_t *n_scal) {
  /* the four loops of precode_matrix_apply_sched: lib/precode.c:23-32 */
  size_t ax = 0, sc = 0;
  long m0 = S->marks[0], m1 = S->marks[1], n = (long)S->nops;
  for (long k = 0; k < m1; k++) S->ops[k].beta ? ax++ : sc++;
  for (long k = m0; k >= 0; k--) S->ops[k].beta ? ax++ : sc++;
  for (long k = m1; k < n; k++) S->ops[k].beta ? ax++ : sc++;
  for (long k = 0; k <= m0; k++) S->ops[k].beta ? ax++ : sc++;
  if (n_axpy) *n_axpy = ax;
  if (n_scal) *n_scal = sc;
  return ax + sc;
}

static void swap_rows(uint8_t *D, size_t pitch, size_t T, int a, int b) {
  if (a == b) return; /* oswaprow: oblas_avx.c:17-31 */
  uint8_t *pa = D + (size_t)a * pitch, *pb = D + (size_t)b * pitch;
  for (size_t k = 0; k < T; k++) {
    uint8_t t = pa[k];
    pa[k] = pb[k];
    pb[k] = t;
  }
}

static void permute_rows(uint8_t *D, size_t pitch, size_t T, int *Pm, int n) {
  /* precode_matrix_permute: lib/precode.c:3-13 */
  for (int i = 0; i < n; i++) {
    int at = i;
    while (Pm[at] >= 0) {
      swap_rows(D, pitch, T, i, Pm[at]);
      int nx = Pm[at];
      Pm[at] = -1;
      at = nx;
    }
  }
}

void orc_intermediate(const orc_sched *S, uint8_t *D, size_t pitch, size_t T) {
  /* precode_matrix_apply_sched + precode_matrix_intermediate:
   * lib/precode.c:23-32,379-389 */
  long m0 = S->marks[0], m1 = S->marks[1], n = (long)S->nops;
  for (long k = 0; k < m1; k++) apply_one(D, pitch, T, S->ops[k]);
  for (long k = m0; k >= 0; k--) apply_one(D, pitch, T, S->ops[k]);
  for (long k = m1; k < n; k++) apply_one(D, pitch, T, S->ops[k]);
  for (long k = 0; k <= m0; k++) apply_one(D, pitch, T, S->ops[k]);
  int *rm = malloc(sizeof(int) * (size_t)S->rows);
  int *cm = malloc(sizeof(int) * (size_t)S->cols);
  memcpy(rm, S->di, sizeof(int) * (size_t)S->rows);
  memcpy(cm, S->c, sizeof(int) * (size_t)S->cols);
  permute_rows(D, pitch, T, rm, S->rows);
  permute_rows(D, pitch, T, cm, S->cols);
  free(rm);
  free(cm);
}

uint64_t orc_fnv1a64(const uint8_t *p, size_t n) { /* hash used by the KAT fixtures */
  uint64_t h = 14695981039346656037ULL;
  for (size_t k = 0; k < n; k++) h = (h ^ p[k]) * 1099511628211ULL;
  return h;
}

/* ------------------------------------------------------ whole-block paths */
int orc_encode_block(int K, size_t T, const uint8_t *src, uint8_t *C_out,
                     size_t pitch, size_t *nops, size_t *n_applied) {
  /* nanorq_generate_symbols: lib/nanorq.c:206-232 (load_symbol_matrix :175-182) */
  orc_params P;
  if (orc_params_init(K, &P)) return -1;
  memset(C_out, 0, (size_t)P.L * pitch);
  for (int e = 0; e < K; e++)
    memcpy(C_out + (size_t)(P.S + P.H + e) * pitch, src + (size_t)e * T, T);
  uint32_t *isi = malloc(sizeof(uint32_t) * (size_t)P.Kprime);
  for (int k = 0; k < P.Kprime; k++) isi[k] = (uint32_t)k;
  int st = 0;
  orc_sched *S = orc_invert(&P, 0, isi, &st);
  free(isi);
  if (!S) return st;
  if (nops) *nops = S->nops;
  if (n_applied) *n_applied = orc_applied_ops(S, NULL, NULL);
  orc_intermediate(S, C_out, pitch, T);
  orc_sched_free(S);
  return 0;
}

int orc_decode_block(int K, size_t T, const uint32_t *esis, const uint8_t *syms,
                     size_t n, uint8_t *out, uint8_t *C_out, size_t pitch,
                     size_t *nops, size_t *n_applied) {
  /* nanorq_decoder_add_symbol + nanorq_repair_block:
   * lib/nanorq.c:478-509, 527-565, 591-631 */
  orc_params P;
  if (orc_params_init(K, &P)) return -1;
  uint8_t *have = calloc((size_t)K, 1);
  size_t nrep = 0, *rep = malloc(sizeof(size_t) * (n ? n : 1));
  uint8_t *seen_rep = NULL;
  uint32_t max_esi = 0;
  for (size_t k = 0; k < n; k++)
    if (esis[k] > max_esi) max_esi = esis[k];
  seen_rep = calloc((size_t)max_esi + 2, 1);
  int gaps = K;
  for (size_t k = 0; k < n; k++) { /* add_symbol: IGN once complete, DUP skipped */
    if (gaps == 0) break;
    uint32_t e = esis[k];
    if (e < (uint32_t)K) {
      if (have[e]) continue;
      have[e] = 1;
      gaps--;
      memcpy(out + (size_t)e * T, syms + k * T, T);
    } else {
      if (seen_rep[e]) continue;
      seen_rep[e] = 1;
      rep[nrep++] = k;
    }
  }
  free(seen_rep);
  int rc = 0;
  if (gaps == 0) goto done;
  if (nrep < (size_t)gaps) {
    rc = 1;
    goto done;
  }
  {
    int overhead = (int)nrep - gaps;
    size_t rows = (size_t)P.L + (size_t)overhead;
    uint8_t *D = calloc(rows * pitch, 1);
    uint32_t *isi = malloc(sizeof(uint32_t) * ((size_t)P.Kprime + (size_t)overhead));
    size_t ri = 0;
    uint32_t pad = (uint32_t)(P.Kprime - K);
    for (int e = 0; e < P.Kprime; e++) isi[e] = (uint32_t)e;
    for (int e = 0; e < K; e++) { /* fill_symbol_matrix_gaps :549-565 */
      uint8_t *row = D + (size_t)(P.S + P.H + e) * pitch;
      if (have[e]) {
        memcpy(row, out + (size_t)e * T, T);
      } else {
        memcpy(row, syms + rep[ri] * T, T);
        isi[e] = esis[rep[ri]] + pad; /* patch_precode_matrix :527-547 */
        ri++;
      }
    }
    for (int x = 0; x < overhead; x++, ri++) {
      memcpy(D + ((size_t)P.L + (size_t)x) * pitch, syms + rep[ri] * T, T);
      isi[P.Kprime + x] = esis[rep[ri]] + pad;
    }
    int st = 0;
    orc_sched *S = orc_invert(&P, overhead, isi, &st);
    free(isi);
    if (!S) {
      free(D);
      rc = st;
      goto done;
    }
    if (nops) *nops = S->nops;
    if (n_applied) *n_applied = orc_applied_ops(S, NULL, NULL);
    orc_intermediate(S, D, pitch, T);
    orc_sched_free(S);
    for (int e = 0; e < K; e++) /* decode_repair_rows :567-577 */
      if (!have[e]) orc_lt_row(&P, D, pitch, (uint32_t)e, out + (size_t)e * T, T);
    if (C_out) memcpy(C_out, D, (size_t)P.L * pitch);
    free(D);
  }
done:
  free(have);
  free(rep);
  return rc;
}
