/* rqb_planner.c -- builds the device solve program for one source block.
 *
 * What the reference does in precode_matrix_gen + precode_matrix_invert
 * (lib/precode.c:90-377) -- generate the sparse constraint matrix A, peel it,
 * eliminate the dense remainder, record row operations -- is re-designed here
 * for a device that holds a column slice of every row in shared memory:
 *
 *   1. build A (LDPC + LT rows; HDPC rows are handled in closed form);
 *   2. peel: order i (row, column) pairs so the peeled part X is unit lower
 *      triangular, inactivating columns when only degree-2 rows are left
 *      (same idea as precode_matrix_precond, lib/precode.c:176-203);
 *   3. bit-matrix work on the host only (never on symbol data):
 *        G      = X^-1 * U_top            (i x u bits)
 *        Schur  = U_low - X_low * G       (binary rows: bits; HDPC rows: GF(256),
 *                                          via the alpha-recurrence of make_HDPC,
 *                                          lib/precode.c:60-83, in O((K'+S)*u))
 *      then a Gauss-Jordan of the small Schur system, binary rows first, so the
 *      inactive symbols z become explicit linear combinations of residual rows;
 *   4. emit gather tasks (rqb_program.h):
 *        A  Y      = X^-1 b_top           sparse forward substitution, by levels
 *        B  r_low  = b_low ^ X_low*Y      + HDPC rows through HORNER chunk scans
 *        C  z      = (Schur)^-1 r_low     3-4 dense levels
 *        D  b_top' = b_top ^ U_top*z      (b_top re-read from the input rows)
 *        E  x      = X^-1 b_top'          same levels as A
 *        O  outputs: C[] in RFC order and/or LT combinations of it.
 *
 * The intermediate symbols are the unique solution of A*C = D when rank(A) = L,
 * so this factorisation yields the same bytes as the reference's op sequence;
 * rank < L is reported exactly when the reference's elimination would fail.
 */
#define _POSIX_C_SOURCE 200809L
#include "rqb_planner.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rfc6330_tables.h"
#include "rqb_gf256.h"
#include "rqb_program.h"

/* ------------------------------------------------------------------ utils */
static rqb_gf_tables GF;
static uint64_t SPREAD[256]; /* byte -> 8 bytes holding its bits as 0/1 */
static volatile int tables_ready;

static void tables_init(void) {
  if (tables_ready) return;
  rqb_gf_build(&GF);
  for (int b = 0; b < 256; b++) {
    uint64_t v = 0;
    for (int k = 0; k < 8; k++)
      if (b >> k & 1) v |= (uint64_t)1 << (8 * k);
    SPREAD[b] = v;
  }
  __sync_synchronize();
  tables_ready = 1;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int rqb_params_init(int K, rqb_params *P) {
  if (K < 1 || K > RQB_MAX_K) return -1;
  int lo = 0, hi = RQB_NUM_KPRIME - 1;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (rqb_kprime_table[mid].kprime >= K)
      hi = mid;
    else
      lo = mid + 1;
  }
  const rqb_kprime_row *r = &rqb_kprime_table[lo];
  P->Kprime = r->kprime;
  P->J = r->j;
  P->S = r->s;
  P->H = r->h;
  P->W = r->w;
  P->L = P->Kprime + P->S + P->H;
  P->P = P->L - P->W;
  P->U = P->P - P->H;
  P->B = P->W - P->S;
  int p1 = P->P;
  for (;; p1++) {
    int prime = p1 >= 2;
    for (int q = 2; q * q <= p1 && prime; q++) prime = (p1 % q) != 0;
    if (prime) break;
  }
  P->P1 = p1;
  return 0;
}

int rqb_host_lt_indices(const rqb_params *P, uint32_t X, uint32_t *out) {
  return rqb_lt_indices(P, rqb_rand_v, rqb_degree_cdf, X, out);
}

/* ------------------------------------------------------ program builder */
typedef struct {
  rqb_task *tasks;
  size_t nt, ct;
  uint8_t *src;
  size_t ns, cs;
  uint8_t *pages;
  size_t npages, cpages;
  uint32_t cur; /* write offset inside the open page, 0 = none open */
  uint32_t levels_in_page;
  size_t tot_levels, tot_tasks, tot_srcs, tot_gf, tot_horner;
  int error;
} progbuf;

static void pb_task(progbuf *pb, int kind, uint32_t dst, uint32_t arg, uint32_t nsrc,
                    const void *srcs, size_t esz) {
  if (pb->nt == pb->ct) {
    pb->ct = pb->ct ? pb->ct * 2 : 1024;
    pb->tasks = realloc(pb->tasks, pb->ct * sizeof(rqb_task));
  }
  size_t bytes = ((size_t)nsrc * esz + 7) & ~(size_t)7;
  if (pb->ns + bytes > pb->cs) {
    pb->cs = (pb->cs ? pb->cs * 2 : 65536) + bytes;
    pb->src = realloc(pb->src, pb->cs);
  }
  if (nsrc > 0xFFFF || dst > 0xFFFF) pb->error = 1;
  rqb_task *t = &pb->tasks[pb->nt++];
  memset(t, 0, sizeof(*t));
  t->src_off = (uint32_t)pb->ns;
  t->arg = arg;
  t->nsrc = (uint16_t)nsrc;
  t->dst = (uint16_t)dst;
  t->kind = (uint8_t)kind;
  if (nsrc) memcpy(pb->src + pb->ns, srcs, (size_t)nsrc * esz);
  if (bytes > (size_t)nsrc * esz) memset(pb->src + pb->ns + (size_t)nsrc * esz, 0, bytes - (size_t)nsrc * esz);
  pb->ns += bytes;
  if (kind == RQB_T_GF_SET || kind == RQB_T_GF_ACC)
    pb->tot_gf += nsrc;
  else if (kind == RQB_T_HORNER)
    pb->tot_horner += nsrc;
  else
    pb->tot_srcs += nsrc;
}

static size_t task_src_bytes(const rqb_task *t) {
  size_t esz = (t->kind == RQB_T_GF_SET || t->kind == RQB_T_GF_ACC || t->kind == RQB_T_HORNER) ? 4 : 2;
  return ((size_t)t->nsrc * esz + 7) & ~(size_t)7;
}

static void pb_close_page(progbuf *pb) {
  if (!pb->cur) return;
  rqb_page_hdr *h = (rqb_page_hdr *)(pb->pages + (pb->npages - 1) * RQB_PAGE_BYTES);
  h->n_levels = pb->levels_in_page;
  pb->cur = 0;
  pb->levels_in_page = 0;
}

static void pb_open_page(progbuf *pb) {
  if (pb->npages == pb->cpages) {
    pb->cpages = pb->cpages ? pb->cpages * 2 : 64;
    pb->pages = realloc(pb->pages, pb->cpages * RQB_PAGE_BYTES);
  }
  memset(pb->pages + pb->npages * RQB_PAGE_BYTES, 0, RQB_PAGE_BYTES);
  pb->npages++;
  pb->cur = sizeof(rqb_page_hdr);
  pb->levels_in_page = 0;
}

static int task_cmp(const void *a, const void *b) {
  const rqb_task *x = a, *y = b;
  if (x->kind != y->kind) return (int)x->kind - (int)y->kind;
  return (int)y->nsrc - (int)x->nsrc;
}

/* close the current level: pack its tasks into pages (splitting when a page
 * fills up; splitting a level is always legal, its tasks are independent) */
static void pb_level_end(progbuf *pb) {
  if (!pb->nt) return;
  qsort(pb->tasks, pb->nt, sizeof(rqb_task), task_cmp);
  size_t idx = 0;
  while (idx < pb->nt) {
    if (!pb->cur) pb_open_page(pb);
    size_t avail = RQB_PAGE_BYTES - pb->cur, need = sizeof(rqb_level_hdr), j = idx;
    while (j < pb->nt) {
      size_t add = sizeof(rqb_task) + task_src_bytes(&pb->tasks[j]);
      if (need + add > avail) break;
      need += add;
      j++;
    }
    if (j == idx) {
      if (pb->cur == sizeof(rqb_page_hdr)) { /* a single task larger than a page */
        pb->error = 1;
        return;
      }
      pb_close_page(pb);
      continue;
    }
    need = (need + 15) & ~(size_t)15; /* levels start 16-byte aligned (uint4 loads); avail is a multiple of 16 */
    uint8_t *page = pb->pages + (pb->npages - 1) * RQB_PAGE_BYTES;
    rqb_level_hdr *lh = (rqb_level_hdr *)(page + pb->cur);
    size_t n = j - idx;
    lh->n_tasks = (uint32_t)n;
    lh->next_off = (uint32_t)(pb->cur + need);
    rqb_task *dst = (rqb_task *)(page + pb->cur + sizeof(rqb_level_hdr));
    uint32_t soff = (uint32_t)(pb->cur + sizeof(rqb_level_hdr) + n * sizeof(rqb_task));
    for (size_t k = 0; k < n; k++) {
      rqb_task t = pb->tasks[idx + k];
      size_t sb = task_src_bytes(&t);
      memcpy(page + soff, pb->src + t.src_off, sb);
      t.src_off = soff;
      soff += (uint32_t)sb;
      dst[k] = t;
    }
    pb->cur += (uint32_t)need;
    pb->levels_in_page++;
    pb->tot_levels++;
    pb->tot_tasks += n;
    idx = j;
    if (RQB_PAGE_BYTES - pb->cur < sizeof(rqb_level_hdr) + sizeof(rqb_task) + 8) pb_close_page(pb);
  }
  pb->nt = 0;
  pb->ns = 0;
}

/* ------------------------------------------------------------ bit helpers */
static inline void bits_xor(uint64_t *a, const uint64_t *b, int w) {
  for (int k = 0; k < w; k++) a[k] ^= b[k];
}
static inline int bit_get(const uint64_t *a, int k) { return (int)(a[k >> 6] >> (k & 63)) & 1; }
static inline void bit_flip(uint64_t *a, int k) { a[k >> 6] ^= (uint64_t)1 << (k & 63); }

/* packed-byte alpha multiply (x -> 2x in GF(256)) on 8 bytes at once */
static inline uint64_t xtime8(uint64_t x) {
  return ((x & 0x7f7f7f7f7f7f7f7fULL) << 1) ^ (((x >> 7) & 0x0101010101010101ULL) * 0x1d);
}

/* ---------------------------------------------------------------- planner */
#define FREE_ALL()                                                                        \
  do {                                                                                    \
    free(rptr); free(cidx); free(cptr); free(ridx); free(deg); free(col_state);           \
    free(col_pos); free(col_t); free(row_pos); free(prow); free(pcol); free(ucol);        \
    free(stk1); free(stk2); free(dptr); free(didx); free(level); free(G); free(lowrows);  \
    free(Sb); free(Tb); free(xptr); free(xidx); free(Sh); free(hb1); free(hb2);           \
    free(pivrow); free(pivcol_of_row); free(freecols); free(Q); free(TQ); free(tmp16);    \
    free(tmp32); free(lvl_cnt); free(lvl_ord); free(cslot); free(ybuf); free(accbuf);     \
    free(pb.tasks); free(pb.src);                                                         \
  } while (0)

int rqb_plan_build(const rqb_plan_request *req, rqb_plan **out) {
  tables_init();
  *out = NULL;
  rqb_params P;
  if (rqb_params_init(req->K, &P) || req->overhead < 0) return -1;
  const int S = P.S, H = P.H, W = P.W, L = P.L, Kp = P.Kprime, B = P.B;
  const int oh = req->overhead, R = L + oh, n = Kp + S, nlt = Kp + oh;
  double t0 = now_s();

  int *rptr = NULL, *cidx = NULL, *cptr = NULL, *ridx = NULL, *deg = NULL;
  uint8_t *col_state = NULL;
  int *col_pos = NULL, *col_t = NULL, *row_pos = NULL, *prow = NULL, *pcol = NULL, *ucol = NULL;
  int *stk1 = NULL, *stk2 = NULL, *dptr = NULL, *didx = NULL, *level = NULL, *lowrows = NULL;
  uint64_t *G = NULL, *Sb = NULL, *Tb = NULL, *ybuf = NULL, *accbuf = NULL;
  int *xptr = NULL, *xidx = NULL;
  uint8_t *Sh = NULL, *hb1 = NULL, *hb2 = NULL, *Q = NULL, *TQ = NULL;
  int *pivrow = NULL, *pivcol_of_row = NULL, *freecols = NULL, *lvl_cnt = NULL, *lvl_ord = NULL;
  uint16_t *tmp16 = NULL, *cslot = NULL;
  uint32_t *tmp32 = NULL;
  progbuf pb;
  memset(&pb, 0, sizeof(pb));
  rqb_plan *plan = NULL;
  int rc = 0;

  /* ---- 1. sparse matrix A, rows: [0,S) LDPC, [S,S+H) HDPC (kept empty, closed form),
   *         [S+H, R) LT rows.  Same contents as precode_matrix_gen (+patching). */
  rptr = calloc((size_t)R + 1, sizeof(int));
  {
    for (int col = 0; col < B; col++) {
      int sub = col / S;
      rptr[1 + col % S]++;
      rptr[1 + (col + sub + 1) % S]++;
      rptr[1 + (col + 2 * (sub + 1)) % S]++;
    }
    for (int r = 0; r < S; r++) rptr[1 + r] += 3;
  }
  size_t cap = (size_t)3 * B + 3 * (size_t)S + (size_t)RQB_MAX_LT_DEGREE * (size_t)nlt + 16;
  cidx = malloc(cap * sizeof(int));
  {
    int acc = 0;
    for (int r = 0; r < S; r++) {
      int c = rptr[1 + r];
      rptr[r] = acc;
      acc += c;
    }
    for (int r = S; r <= S + H; r++) rptr[r] = acc;
    int *cur = malloc(sizeof(int) * (size_t)S);
    memcpy(cur, rptr, sizeof(int) * (size_t)S);
    for (int col = 0; col < B; col++) {
      int sub = col / S;
      cidx[cur[col % S]++] = col;
      cidx[cur[(col + sub + 1) % S]++] = col;
      cidx[cur[(col + 2 * (sub + 1)) % S]++] = col;
    }
    for (int r = 0; r < S; r++) {
      cidx[cur[r]++] = B + r;
      cidx[cur[r]++] = W + r % P.P;
      cidx[cur[r]++] = W + (r + 1) % P.P;
    }
    free(cur);
    uint32_t idx[RQB_MAX_LT_DEGREE];
    for (int k = 0; k < nlt; k++) {
      int cnt = rqb_host_lt_indices(&P, req->isi[k], idx);
      for (int q = 0; q < cnt; q++) cidx[acc + q] = (int)idx[q];
      acc += cnt;
      rptr[S + H + k + 1] = acc;
    }
  }
  const int nnz = rptr[R];
  /* column lists */
  cptr = calloc((size_t)L + 1, sizeof(int));
  ridx = malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
  for (int k = 0; k < nnz; k++) cptr[cidx[k] + 1]++;
  for (int c = 0; c < L; c++) cptr[c + 1] += cptr[c];
  {
    int *cur = malloc(sizeof(int) * (size_t)L);
    memcpy(cur, cptr, sizeof(int) * (size_t)L);
    for (int r = 0; r < R; r++)
      for (int k = rptr[r]; k < rptr[r + 1]; k++) ridx[cur[cidx[k]]++] = r;
    free(cur);
  }
  double t1 = now_s();

  /* ---- 2. peeling */
  deg = calloc((size_t)R, sizeof(int));
  col_state = calloc((size_t)L, 1); /* 0 active, 1 peeled, 2 inactive */
  col_pos = malloc(sizeof(int) * (size_t)L);
  col_t = malloc(sizeof(int) * (size_t)L);
  row_pos = malloc(sizeof(int) * (size_t)R);
  prow = malloc(sizeof(int) * (size_t)L);
  pcol = malloc(sizeof(int) * (size_t)L);
  ucol = malloc(sizeof(int) * (size_t)L);
  stk1 = malloc(sizeof(int) * ((size_t)nnz + (size_t)R + 8));
  stk2 = malloc(sizeof(int) * ((size_t)nnz + (size_t)R + 8));
  int n1 = 0, n2 = 0, ni = 0, nu = 0;
  for (int r = 0; r < R; r++) row_pos[r] = -1;
  for (int c = 0; c < L; c++) col_pos[c] = col_t[c] = -1;
  for (int c = W; c < L; c++) { /* the P permanently inactive columns */
    col_state[c] = 2;
    col_t[c] = nu;
    ucol[nu++] = c;
  }
  for (int r = 0; r < R; r++)
    for (int k = rptr[r]; k < rptr[r + 1]; k++) deg[r] += (cidx[k] < W);
  for (int r = S + H; r < R; r++) {
    if (deg[r] == 1) stk1[n1++] = r;
    if (deg[r] == 2) stk2[n2++] = r;
  }
  for (int r = 0; r < S; r++) {
    if (deg[r] == 1) stk1[n1++] = r;
    if (deg[r] == 2) stk2[n2++] = r;
  }
  while (ni + nu < L) {
    int r = -1;
    while (n1 > 0) {
      int cand = stk1[--n1];
      if (row_pos[cand] < 0 && deg[cand] == 1) {
        r = cand;
        break;
      }
    }
    if (r < 0)
      while (n2 > 0) {
        int cand = stk2[--n2];
        if (row_pos[cand] < 0 && deg[cand] == 2) {
          r = cand;
          break;
        }
      }
    if (r < 0) break;
    int c0 = -1, c1 = -1;
    for (int k = rptr[r]; k < rptr[r + 1]; k++) {
      int c = cidx[k];
      if (col_state[c] == 0) {
        if (c0 < 0)
          c0 = c;
        else
          c1 = c;
      }
    }
    if (c1 >= 0 && cptr[c1 + 1] - cptr[c1] < cptr[c0 + 1] - cptr[c0]) {
      int t = c0; /* inactivate the heavier column, pivot on the lighter one */
      c0 = c1;
      c1 = t;
    }
    row_pos[r] = ni;
    prow[ni] = r;
    pcol[ni] = c0;
    col_state[c0] = 1;
    col_pos[c0] = ni;
    ni++;
    for (int k = cptr[c0]; k < cptr[c0 + 1]; k++) {
      int rr = ridx[k], dg = --deg[rr];
      if (dg == 1) stk1[n1++] = rr;
      else if (dg == 2) stk2[n2++] = rr;
    }
    if (c1 >= 0) {
      col_state[c1] = 2;
      col_t[c1] = nu;
      ucol[nu++] = c1;
      for (int k = cptr[c1]; k < cptr[c1 + 1]; k++) {
        int rr = ridx[k], dg = --deg[rr];
        if (dg == 1) stk1[n1++] = rr;
        else if (dg == 2) stk2[n2++] = rr;
      }
    }
  }
  for (int c = 0; c < W; c++) /* whatever could not be peeled is inactive too */
    if (col_state[c] == 0) {
      col_state[c] = 2;
      col_t[c] = nu;
      ucol[nu++] = c;
    }
  const int I = ni, U = nu, uw = (U + 63) / 64;
  const int nb = R - H - I;
  if (I + U != L || nb < 0) {
    rc = -2;
    goto fail;
  }
  double t2 = now_s();

  /* ---- 3a. dependencies, levels and G = X^-1 U_top (bits) */
  dptr = malloc(sizeof(int) * ((size_t)I + 1));
  didx = malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
  level = malloc(sizeof(int) * ((size_t)I + 1));
  G = calloc((size_t)(I ? I : 1) * uw, sizeof(uint64_t));
  int maxlevel = -1;
  {
    int nd = 0;
    for (int p = 0; p < I; p++) {
      int r = prow[p], lv = 0;
      uint64_t *g = G + (size_t)p * uw;
      dptr[p] = nd;
      for (int k = rptr[r]; k < rptr[r + 1]; k++) {
        int c = cidx[k];
        if (col_state[c] == 2) {
          bit_flip(g, col_t[c]);
        } else if (c != pcol[p]) {
          int q = col_pos[c];
          if (q >= p) { /* cannot happen: would contradict the peeling invariant */
            rc = -3;
            goto fail;
          }
          didx[nd++] = q;
          if (level[q] + 1 > lv) lv = level[q] + 1;
          bits_xor(g, G + (size_t)q * uw, uw);
        }
      }
      level[p] = lv;
      if (lv > maxlevel) maxlevel = lv;
    }
    dptr[I] = nd;
  }

  /* ---- 3b. residual binary rows: Schur bits and their X-part source lists */
  lowrows = malloc(sizeof(int) * (size_t)(nb ? nb : 1));
  {
    int m = 0;
    for (int r = S + H; r < R; r++)
      if (row_pos[r] < 0) lowrows[m++] = r;
    for (int r = 0; r < S; r++)
      if (row_pos[r] < 0) lowrows[m++] = r;
  }
  const int nbw = (nb + 63) / 64;
  Sb = calloc((size_t)(nb ? nb : 1) * uw, sizeof(uint64_t));
  Tb = calloc((size_t)(nb ? nb : 1) * (nbw ? nbw : 1), sizeof(uint64_t));
  xptr = malloc(sizeof(int) * ((size_t)nb + 1));
  xidx = malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
  {
    int nx = 0;
    for (int m = 0; m < nb; m++) {
      int r = lowrows[m];
      uint64_t *s = Sb + (size_t)m * uw;
      xptr[m] = nx;
      for (int k = rptr[r]; k < rptr[r + 1]; k++) {
        int c = cidx[k];
        if (col_state[c] == 2) {
          bit_flip(s, col_t[c]);
        } else {
          int q = col_pos[c];
          xidx[nx++] = q;
          bits_xor(s, G + (size_t)q * uw, uw);
        }
      }
      bit_flip(Tb + (size_t)m * nbw, m);
    }
    xptr[nb] = nx;
  }

  /* ---- 3c. HDPC Schur rows (H x U bytes) by the alpha recurrence.
   * HDPC[:,j] = alpha*HDPC[:,j+1] ^ e_b1(j) ^ e_b2(j), last column alpha^h
   * (lib/precode.c:60-83)  =>  sum_j HDPC[h][j] v_j = alpha^h y_{n-1} ^ sum_{j<=n-2, h in b(j)} y_j
   * with y_j = alpha*y_{j-1} ^ v_j.  Here v_j is the u-vector of column j: row q of G if the
   * column is peeled at q (its symbol is Y_q ^ G_q z), the unit vector if inactive. */
  hb1 = malloc((size_t)n);
  hb2 = malloc((size_t)n);
  for (int j = 0; j + 1 < n; j++) {
    uint32_t b1 = rqb_rand(rqb_rand_v, (uint32_t)j + 1, 6, (uint32_t)H);
    uint32_t b2 = (b1 + rqb_rand(rqb_rand_v, (uint32_t)j + 1, 7, (uint32_t)H - 1) + 1) % (uint32_t)H;
    hb1[j] = (uint8_t)b1;
    hb2[j] = (uint8_t)b2;
  }
  const int uq = uw * 8; /* u64 words per byte-row of padded width 64*uw */
  Sh = calloc((size_t)H * (size_t)uq * 8, 1);
  ybuf = calloc((size_t)uq, 8);
  accbuf = calloc((size_t)H * (size_t)uq, 8);
  {
    for (int j = 0; j < n; j++) {
      for (int k = 0; k < uq; k++) ybuf[k] = xtime8(ybuf[k]);
      if (col_state[j] == 1) {
        const uint8_t *g = (const uint8_t *)(G + (size_t)col_pos[j] * uw);
        for (int k = 0; k < uq; k++) ybuf[k] ^= SPREAD[g[k]];
      } else {
        ((uint8_t *)ybuf)[col_t[j]] ^= 1;
      }
      if (j + 1 < n) {
        uint64_t *a1 = accbuf + (size_t)hb1[j] * uq, *a2 = accbuf + (size_t)hb2[j] * uq;
        for (int k = 0; k < uq; k++) {
          a1[k] ^= ybuf[k];
          a2[k] ^= ybuf[k];
        }
      }
    }
    for (int h = 0; h < H; h++) {
      uint8_t *row = Sh + (size_t)h * (size_t)uq * 8;
      const uint8_t *a = (const uint8_t *)(accbuf + (size_t)h * uq), *y = (const uint8_t *)ybuf;
      uint8_t ah = rqb_gf_pow2(&GF, h);
      for (int t = 0; t < U; t++) row[t] = a[t] ^ rqb_gf_mul(&GF, ah, y[t]);
      row[col_t[n + h]] ^= 1; /* the identity block I_H */
    }
  }

  /* ---- 3d. Gauss-Jordan over GF(2) on the binary Schur rows, transformation tracked */
  pivrow = malloc(sizeof(int) * (size_t)(U ? U : 1));       /* column t -> local row or -1 */
  pivcol_of_row = malloc(sizeof(int) * (size_t)(nb ? nb : 1));
  freecols = malloc(sizeof(int) * (size_t)(U ? U : 1));
  int rho = 0, nfree = 0;
  for (int m = 0; m < nb; m++) pivcol_of_row[m] = -1;
  for (int t = 0; t < U; t++) {
    int pr = -1;
    for (int m = 0; m < nb; m++)
      if (pivcol_of_row[m] < 0 && bit_get(Sb + (size_t)m * uw, t)) {
        pr = m;
        break;
      }
    pivrow[t] = pr;
    if (pr < 0) {
      freecols[nfree++] = t;
      continue;
    }
    pivcol_of_row[pr] = t;
    rho++;
    const uint64_t *ps = Sb + (size_t)pr * uw, *pt = Tb + (size_t)pr * nbw;
    for (int m = 0; m < nb; m++)
      if (m != pr && bit_get(Sb + (size_t)m * uw, t)) {
        bits_xor(Sb + (size_t)m * uw, ps, uw);
        bits_xor(Tb + (size_t)m * nbw, pt, nbw);
      }
  }
  if (nfree > H) {
    rc = 1;
    goto fail;
  }
  /* ---- 3e. HDPC rows: eliminate pivot columns (beta = Sh[h][t]), then solve the
   *          H x nfree system Q over GF(256) with a tracked transformation TQ (H x H) */
  Q = calloc((size_t)H * (size_t)(nfree ? nfree : 1), 1);
  TQ = calloc((size_t)H * (size_t)H, 1);
  for (int h = 0; h < H; h++) {
    const uint8_t *row = Sh + (size_t)h * (size_t)uq * 8;
    for (int f = 0; f < nfree; f++) Q[h * nfree + f] = row[freecols[f]];
    for (int t = 0; t < U; t++) {
      uint8_t beta = row[t];
      if (!beta || pivrow[t] < 0) continue;
      const uint64_t *ps = Sb + (size_t)pivrow[t] * uw;
      for (int f = 0; f < nfree; f++)
        if (bit_get(ps, freecols[f])) Q[h * nfree + f] ^= beta;
    }
    TQ[h * H + h] = 1;
  }
  int qrow_of_f[16];
  {
    uint8_t used[16] = {0};
    for (int f = 0; f < nfree; f++) {
      int pr = -1;
      for (int h = 0; h < H; h++)
        if (!used[h] && Q[h * nfree + f]) {
          pr = h;
          break;
        }
      if (pr < 0) {
        rc = 1; /* rank(A) < L */
        goto fail;
      }
      used[pr] = 1;
      qrow_of_f[f] = pr;
      uint8_t inv = GF.inv[Q[pr * nfree + f]];
      for (int k = 0; k < nfree; k++) Q[pr * nfree + k] = rqb_gf_mul(&GF, Q[pr * nfree + k], inv);
      for (int k = 0; k < H; k++) TQ[pr * H + k] = rqb_gf_mul(&GF, TQ[pr * H + k], inv);
      for (int h = 0; h < H; h++) {
        uint8_t b = Q[h * nfree + f];
        if (h == pr || !b) continue;
        for (int k = 0; k < nfree; k++) Q[h * nfree + k] ^= rqb_gf_mul(&GF, b, Q[pr * nfree + k]);
        for (int k = 0; k < H; k++) TQ[h * H + k] ^= rqb_gf_mul(&GF, b, TQ[pr * H + k]);
      }
    }
  }
  double t3 = now_s();

  /* ---- 4. emit the program.  Slots: matrix row r -> slot r; scratch after R. */
  const int NC = (n + 1023) / 1024 > 32 ? (n + 1023) / 1024 : (n < 64 ? 1 : 32); /* HORNER chunks */
  const uint32_t RP = (uint32_t)R;                 /* r'_t, one per pivot column (indexed by t) */
  const uint32_t Z = RP + (uint32_t)U;             /* z_t                                         */
  const uint32_t HS = Z + (uint32_t)U;             /* HORNER scratch: NC * (H+1)                  */
  const uint32_t PS = HS + (uint32_t)NC * (uint32_t)(H + 1); /* GF partial sums               */
  const int GFCH = 8;                              /* GF sources per partial task               */
  const int parts_per_h = (rho + NC + GFCH - 1) / GFCH + 1;
  const uint32_t n_slots = PS + (uint32_t)H * (uint32_t)parts_per_h;
  if (n_slots > RQB_MAX_SLOTS) {
    rc = -4;
    goto fail;
  }
  size_t tmpcap = (size_t)(I > U ? I : U) + (size_t)nb + (size_t)n + 4096;
  tmp16 = malloc(sizeof(uint16_t) * tmpcap);
  tmp32 = malloc(sizeof(uint32_t) * tmpcap);

  /* counting sort of peeled positions by level */
  const int nlev = maxlevel + 1;
  lvl_cnt = calloc((size_t)nlev + 2, sizeof(int));
  lvl_ord = malloc(sizeof(int) * (size_t)(I ? I : 1));
  for (int p = 0; p < I; p++) lvl_cnt[level[p] + 1]++;
  for (int l = 0; l < nlev; l++) lvl_cnt[l + 1] += lvl_cnt[l];
  {
    int *cur = malloc(sizeof(int) * ((size_t)nlev + 1));
    memcpy(cur, lvl_cnt, sizeof(int) * ((size_t)nlev + 1));
    for (int p = 0; p < I; p++) lvl_ord[cur[level[p]]++] = p;
    free(cur);
  }

  /* A: forward substitution Y = X^-1 b_top (level 0 rows have no sources: nothing to do) */
  for (int l = 1; l < nlev; l++) {
    for (int k = lvl_cnt[l]; k < lvl_cnt[l + 1]; k++) {
      int p = lvl_ord[k], ns = 0;
      for (int e = dptr[p]; e < dptr[p + 1]; e++) tmp16[ns++] = (uint16_t)prow[didx[e]];
      pb_task(&pb, RQB_T_XOR_ACC, (uint32_t)prow[p], 0, (uint32_t)ns, tmp16, 2);
    }
    pb_level_end(&pb);
  }
  /* B: residual rows r_m ^= X_low*Y ; HDPC rows through NC chunk scans over columns 0..n-1 */
  for (int m = 0; m < nb; m++) {
    int ns = 0;
    for (int e = xptr[m]; e < xptr[m + 1]; e++) tmp16[ns++] = (uint16_t)prow[xidx[e]];
    if (ns) pb_task(&pb, RQB_T_XOR_ACC, (uint32_t)lowrows[m], 0, (uint32_t)ns, tmp16, 2);
  }
  int *cs = malloc(sizeof(int) * ((size_t)NC + 1));
  for (int c = 0; c <= NC; c++) cs[c] = (int)((long)n * c / NC);
  for (int c = 0; c < NC; c++) {
    int ns = 0;
    for (int j = cs[c]; j < cs[c + 1]; j++) {
      uint32_t slot = col_state[j] == 1 ? (uint32_t)prow[col_pos[j]] : RQB_SLOT_NONE;
      int fl = (j + 1 < n);
      tmp32[ns++] = RQB_HORNER_ENTRY(slot, fl ? hb1[j] : 0, fl ? hb2[j] : 0, fl);
    }
    pb_task(&pb, RQB_T_HORNER, HS + (uint32_t)c * (uint32_t)(H + 1), (uint32_t)H, (uint32_t)ns, tmp32, 4);
  }
  pb_level_end(&pb);

  /* C1: r'_t = XOR_{m in Tb[pivrow t]} r_m ; HDPC base r_h = XOR_c acc_c[h] */
  for (int t = 0; t < U; t++) {
    if (pivrow[t] < 0) continue;
    const uint64_t *tb = Tb + (size_t)pivrow[t] * nbw;
    int ns = 0;
    for (int m = 0; m < nb; m++)
      if (bit_get(tb, m)) tmp16[ns++] = (uint16_t)lowrows[m];
    pb_task(&pb, RQB_T_XOR_SET, RP + (uint32_t)t, 0, (uint32_t)ns, tmp16, 2);
  }
  for (int h = 0; h < H; h++) {
    for (int c = 0; c < NC; c++) tmp16[c] = (uint16_t)(HS + (uint32_t)c * (uint32_t)(H + 1) + (uint32_t)h);
    pb_task(&pb, RQB_T_XOR_SET, (uint32_t)(S + h), 0, (uint32_t)NC, tmp16, 2);
  }
  pb_level_end(&pb);

  /* C2: r'_h = r_h ^ sum_c Gc[h][c]*yend_c ^ sum_t beta[h][t]*r'_t, as partial GF sums.
   * Gc[h][c'] = alpha^h alpha^(n-e_c') ^ sum_{c>c'} coef[c][h] alpha^(s_c-e_c'),
   * coef[c][h] = sum_{j in chunk c, j<=n-2, h in b(j)} alpha^(j-s_c+1). */
  {
    uint8_t *coef = calloc((size_t)NC * (size_t)H, 1);
    for (int c = 0; c < NC; c++)
      for (int j = cs[c]; j < cs[c + 1] && j + 1 < n; j++) {
        uint8_t a = rqb_gf_pow2(&GF, j - cs[c] + 1);
        coef[c * H + hb1[j]] ^= a;
        coef[c * H + hb2[j]] ^= a;
      }
    for (int h = 0; h < H; h++) {
      const uint8_t *row = Sh + (size_t)h * (size_t)uq * 8;
      int ns = 0;
      for (int c1 = 0; c1 < NC; c1++) {
        uint8_t g = rqb_gf_mul(&GF, rqb_gf_pow2(&GF, h), rqb_gf_pow2(&GF, n - cs[c1 + 1]));
        for (int c = c1 + 1; c < NC; c++)
          g ^= rqb_gf_mul(&GF, coef[c * H + h], rqb_gf_pow2(&GF, cs[c] - cs[c1 + 1]));
        if (g) tmp32[ns++] = (HS + (uint32_t)c1 * (uint32_t)(H + 1) + (uint32_t)H) | ((uint32_t)g << 16);
      }
      for (int t = 0; t < U; t++)
        if (pivrow[t] >= 0 && row[t]) tmp32[ns++] = (RP + (uint32_t)t) | ((uint32_t)row[t] << 16);
      int part = 0;
      for (int o = 0; o < ns; o += GFCH, part++) {
        int cnt = ns - o < GFCH ? ns - o : GFCH;
        pb_task(&pb, RQB_T_GF_SET, PS + (uint32_t)h * (uint32_t)parts_per_h + (uint32_t)part, 0,
                (uint32_t)cnt, tmp32 + o, 4);
      }
      tmp16[h] = (uint16_t)part; /* remember the count for the combine level */
    }
    free(coef);
    uint16_t nparts[16];
    for (int h = 0; h < H; h++) nparts[h] = tmp16[h];
    pb_level_end(&pb);
    for (int h = 0; h < H; h++) {
      for (int q = 0; q < nparts[h]; q++)
        tmp16[q] = (uint16_t)(PS + (uint32_t)h * (uint32_t)parts_per_h + (uint32_t)q);
      if (nparts[h]) pb_task(&pb, RQB_T_XOR_ACC, (uint32_t)(S + h), 0, nparts[h], tmp16, 2);
    }
    pb_level_end(&pb);
  }
  /* C3: z_f = sum_h TQ[qrow(f)][h] * r'_h */
  for (int f = 0; f < nfree; f++) {
    int ns = 0;
    for (int h = 0; h < H; h++) {
      uint8_t b = TQ[qrow_of_f[f] * H + h];
      if (b) tmp32[ns++] = (uint32_t)(S + h) | ((uint32_t)b << 16);
    }
    pb_task(&pb, RQB_T_GF_SET, Z + (uint32_t)freecols[f], 0, (uint32_t)ns, tmp32, 4);
  }
  pb_level_end(&pb);
  /* C4: z_t = r'_t ^ XOR_{f: bit} z_f for pivot columns */
  for (int t = 0; t < U; t++) {
    if (pivrow[t] < 0) continue;
    const uint64_t *ps = Sb + (size_t)pivrow[t] * uw;
    int ns = 0;
    tmp16[ns++] = (uint16_t)(RP + (uint32_t)t);
    for (int f = 0; f < nfree; f++)
      if (bit_get(ps, freecols[f])) tmp16[ns++] = (uint16_t)(Z + (uint32_t)freecols[f]);
    pb_task(&pb, RQB_T_XOR_SET, Z + (uint32_t)t, 0, (uint32_t)ns, tmp16, 2);
  }
  pb_level_end(&pb);

  /* load map (also used by D) */
  plan = calloc(1, sizeof(*plan));
  plan->load_src = malloc(sizeof(uint32_t) * n_slots);
  for (uint32_t s = 0; s < n_slots; s++) plan->load_src[s] = RQB_ROW_NONE;
  for (int k = 0; k < nlt; k++) plan->load_src[S + H + k] = req->in_row[k];

  /* D: b_top' = b_top ^ U_top z   (b_top re-read from the input) */
  for (int p = 0; p < I; p++) {
    int r = prow[p], ns = 0;
    for (int k = rptr[r]; k < rptr[r + 1]; k++)
      if (col_state[cidx[k]] == 2) tmp16[ns++] = (uint16_t)(Z + (uint32_t)col_t[cidx[k]]);
    if (ns == 0 && dptr[p + 1] == dptr[p]) continue; /* x = b: the slot already holds it */
    pb_task(&pb, RQB_T_LOAD_XOR, (uint32_t)r, plan->load_src[r], (uint32_t)ns, tmp16, 2);
  }
  pb_level_end(&pb);
  /* E: x = X^-1 b_top' */
  for (int l = 1; l < nlev; l++) {
    for (int k = lvl_cnt[l]; k < lvl_cnt[l + 1]; k++) {
      int p = lvl_ord[k], ns = 0;
      for (int e = dptr[p]; e < dptr[p + 1]; e++) tmp16[ns++] = (uint16_t)prow[didx[e]];
      pb_task(&pb, RQB_T_XOR_ACC, (uint32_t)prow[p], 0, (uint32_t)ns, tmp16, 2);
    }
    pb_level_end(&pb);
  }
  /* O: outputs.  C[col] sits in the slot of the row that pivoted on col, or in z. */
  cslot = malloc(sizeof(uint16_t) * (size_t)L);
  for (int c = 0; c < L; c++)
    cslot[c] = col_state[c] == 1 ? (uint16_t)prow[col_pos[c]] : (uint16_t)(Z + (uint32_t)col_t[c]);
  if (req->want_c)
    for (int c = 0; c < L; c++) pb_task(&pb, RQB_T_OUT_C, 0, (uint32_t)c, 1, &cslot[c], 2);
  for (int k = 0; k < req->n_out; k++) {
    uint32_t idx[RQB_MAX_LT_DEGREE];
    int cnt = rqb_host_lt_indices(&P, req->out_isi[k], idx);
    for (int q = 0; q < cnt; q++) tmp16[q] = cslot[idx[q]];
    pb_task(&pb, RQB_T_OUT_SYM, 0, (uint32_t)k, (uint32_t)cnt, tmp16, 2);
  }
  pb_level_end(&pb);
  pb_close_page(&pb);
  free(cs);
  if (pb.error) {
    rc = -5;
    goto fail;
  }
  double t4 = now_s();

  plan->P = P;
  plan->K = req->K;
  plan->overhead = oh;
  plan->n_slots = n_slots;
  plan->n_pages = (uint32_t)pb.npages;
  plan->pages = pb.pages;
  pb.pages = NULL;
  plan->n_c_rows = req->want_c ? (uint32_t)L : 0;
  plan->n_out = (uint32_t)req->n_out;
  plan->st.i = I;
  plan->st.u = U;
  plan->st.nb = nb;
  plan->st.rho = rho;
  plan->st.nfree = nfree;
  plan->st.levels_fwd = nlev;
  plan->st.n_levels = (int)pb.tot_levels;
  plan->st.n_tasks = (int)pb.tot_tasks;
  plan->st.n_pages = (int)pb.npages;
  plan->st.n_srcs = pb.tot_srcs;
  plan->st.n_gf_srcs = pb.tot_gf;
  plan->st.n_horner = pb.tot_horner;
  plan->st.nnz = (size_t)nnz;
  plan->st.t_matrix = t1 - t0;
  plan->st.t_peel = t2 - t1;
  plan->st.t_dense = t3 - t2;
  plan->st.t_emit = t4 - t3;
  FREE_ALL();
  *out = plan;
  return 0;

fail:
  if (plan) {
    free(plan->load_src);
    free(plan);
  }
  free(pb.pages);
  FREE_ALL();
  return rc;
}

void rqb_plan_free(rqb_plan *p) {
  if (!p) return;
  free(p->load_src);
  free(p->pages);
  free(p);
}
