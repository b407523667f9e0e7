/* rqb_planner.c -- builds the device solve program for one source block.
 *
 * What the reference does in precode_matrix_gen + precode_matrix_invert
 * (lib/precode.c:90-377) -- generate the sparse constraint matrix A, peel it,
 * eliminate the dense remainder, record row operations -- is re-designed here
 * for a device that replays a levelled gather program on column slices of the
 * block (rqb_program.h):
 *
 *   1. build A (LDPC + LT rows; HDPC rows are handled in closed form);
 *   2. peel: order i (row, column) pairs so the peeled part X is unit lower
 *      triangular, inactivating columns when only degree-2 rows are left
 *      (same idea as precode_matrix_precond, lib/precode.c:176-203), but rows with one
 *      column left are taken breadth-first, which halves the dependency depth of X;
 *   3. bit-matrix work on the host only (never on symbol data):
 *        G      = X^-1 * U_top            (i x u bits)
 *        Schur  = U_low - X_low * G       (binary rows: bits; HDPC rows: GF(256),
 *                                          via the alpha-recurrence of make_HDPC,
 *                                          lib/precode.c:60-83, in O((K'+S)*u))
 *      then a Gauss-Jordan of the small Schur system, binary rows first, so the
 *      inactive symbols z become explicit linear combinations of residual rows;
 *   4. emit gather tasks of at most RQB_MAX_SRCS sources:
 *        A  Y      = X^-1 b_top           sparse forward substitution, by levels;
 *                                         rows with many terms are pre-reduced by
 *                                         partial sums scheduled as early as their
 *                                         inputs exist, off the critical path
 *        B  r_low  = b_low ^ X_low*Y      + alpha-scans of Y in chunks (SCAN tasks)
 *        C  z      = (Schur)^-1 r_low     a few dense levels (reduction trees)
 *        D  b_top' = b_top ^ U_top*z      (b_top re-read from the input rows)
 *        E  x      = X^-1 b_top'          same levels as A
 *        O  outputs: C[] in RFC order and/or LT combinations of it.
 *
 * The intermediate symbols are the unique solution of A*C = D when rank(A) = L,
 * so this factorisation yields the same bytes as the reference's op sequence;
 * rank < L is reported exactly when the reference's elimination would fail.
 *
 * Allocation: every temporary lives in a per-thread scratch arena that is kept
 * between calls, so that decoder threads planning different blocks at wire rate
 * never meet in the allocator (malloc/mmap/page-fault contention made the
 * planner scale negatively with threads).
 */
#define _POSIX_C_SOURCE 200809L
#include "rqb_planner.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rfc6330_tables.h"
#include "rqb_gf256.h"
#include "rqb_device.h"
#include "rqb_program.h"

/* ------------------------------------------------------------------ utils */
static rqb_gf_tables GF;
static uint64_t SPREAD[256]; /* byte -> 8 bytes holding its bits as 0/1 */
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

static void tables_build(void) {
  rqb_gf_build(&GF);
  for (int b = 0; b < 256; b++) {
    uint64_t v = 0;
    for (int k = 0; k < 8; k++)
      if (b >> k & 1) v |= (uint64_t)1 << (8 * k);
    SPREAD[b] = v;
  }
}

#ifdef RQB_PLAN_FINE
double rqb_plan_fine[24];
#define FINE(k) do { double _t = now_s(); rqb_plan_fine[k] += _t - fine_t; fine_t = _t; } while (0)
#else
#define FINE(k) do { } while (0)
#endif
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

/* ------------------------------------------------- per-thread scratch arena */
enum {
  SC_RPTR, SC_CIDX, SC_CPTR, SC_RIDX, SC_CUR, SC_DEG, SC_COLSTATE, SC_COLPOS, SC_COLT, SC_ROWPOS, SC_PROW,
  SC_PCOL, SC_UCOL, SC_STK1, SC_STK2, SC_LEVEL, SC_G, SC_LOWROWS, SC_SB, SC_TB, SC_XPTR, SC_XIDX, SC_SH,
  SC_HB1, SC_HB2, SC_YBUF, SC_ACCBUF, SC_PIVROW, SC_PIVCOL, SC_FREECOLS, SC_Q, SC_TQ, SC_TMP, SC_CSLOT,
  SC_FPTR, SC_FITEMS, SC_PFIRST, SC_PLEVEL, SC_PPTR, SC_PITEMS, SC_CURLOC, SC_TASKS, SC_SRCS, SC_ORDER,
  SC_LVLCNT, SC_CS, SC_COEF, SC_SORT, SC_FRUSED, SC_FRTAB, SC_CINFO, SC_CHFIRST, SC_CHLEVEL, SC_CHPTR,
  SC_CHITEMS, SC_HAS, SC_NEEDED, SC_GC, SC_COUNT
};
typedef struct {
  void *p[SC_COUNT];
  size_t cap[SC_COUNT];
  int oom; /* an allocation failed since the flag was last cleared */
} scratch_t;

static pthread_key_t sc_key;
static pthread_once_t sc_once = PTHREAD_ONCE_INIT;
static void sc_destroy(void *v) {
  scratch_t *sc = v;
  if (!sc) return;
  for (int k = 0; k < SC_COUNT; k++) free(sc->p[k]);
  free(sc);
}
static void sc_make_key(void) { pthread_key_create(&sc_key, sc_destroy); }
static scratch_t *sc_get(void) {
  pthread_once(&sc_once, sc_make_key);
  scratch_t *sc = pthread_getspecific(sc_key);
  if (!sc) {
    sc = calloc(1, sizeof(*sc));
    pthread_setspecific(sc_key, sc);
  }
  return sc;
}
/* a buffer of at least `bytes` (contents undefined unless zero != 0) */
static void *sc_buf(scratch_t *sc, int id, size_t bytes, int zero) {
  if (bytes < 64) bytes = 64;
  if (sc->cap[id] < bytes) {
    free(sc->p[id]);
    size_t cap = bytes + bytes / 2;
    sc->p[id] = malloc(cap);
    sc->cap[id] = sc->p[id] ? cap : 0;
    if (!sc->p[id]) {
      sc->oom = 1;
      return NULL;
    }
  }
  if (zero) memset(sc->p[id], 0, bytes);
  return sc->p[id];
}
/* grow keeping the contents */
static void *sc_grow(scratch_t *sc, int id, size_t bytes) {
  if (sc->cap[id] < bytes) {
    size_t cap = bytes * 2;
    void *np_ = realloc(sc->p[id], cap);
    if (!np_) {
      sc->oom = 1;
      return sc->p[id];
    }
    sc->p[id] = np_;
    sc->cap[id] = cap;
  }
  return sc->p[id];
}

/* ------------------------------------------------------ program builder */
typedef struct {
  uint32_t dst, src_at, level;
  uint32_t extra; /* TAB: src0 */
  uint16_t nsrc;
  uint8_t kind, aux;
} ptask;

typedef struct {
  scratch_t *sc;
  ptask *tasks;
  size_t nt;
  uint32_t *srcs;
  size_t ns;
  uint32_t ws_base; /* arena row of working row 0 */
  uint32_t ws_next; /* next unused working row */
  uint32_t tab_base; /* first row of the XOR tables of the back-substitution */
  uint32_t max_level;
  size_t tot_x, tot_gf, tot_h;
} builder;

static void b_task(builder *b, int kind, uint32_t dst, int aux, uint32_t level, const uint32_t *srcs, uint32_t n) {
  b->tasks = sc_grow(b->sc, SC_TASKS, (b->nt + 1) * sizeof(ptask));
  b->srcs = sc_grow(b->sc, SC_SRCS, (b->ns + n + 1) * sizeof(uint32_t));
  if (b->sc->oom) return; /* rqb_plan_build reports it */
  ptask *t = &b->tasks[b->nt++];
  t->dst = dst;
  t->src_at = (uint32_t)b->ns;
  t->level = level;
  t->nsrc = (uint16_t)n;
  t->kind = (uint8_t)kind;
  t->aux = (uint8_t)aux;
  t->extra = 0;
  if (kind == RQB_T_GF || kind == RQB_T_SCAN2) { /* the top byte carries the multiplier / the two HDPC rows */
    if (n) memcpy(b->srcs + b->ns, srcs, (size_t)n * sizeof(uint32_t));
  } else { /* XOR and SCAN sources are bare row numbers */
    for (uint32_t k = 0; k < n; k++) b->srcs[b->ns + k] = srcs[k] & RQB_REF_MASK;
  }
  b->ns += n;
  if (level > b->max_level) b->max_level = level;
  if (kind == RQB_T_GF)
    b->tot_gf += n;
  else if (kind == RQB_T_SCAN || kind == RQB_T_SCAN2)
    b->tot_h += n;
  else
    b->tot_x += n;
}

/* TAB task: row[dst] = row[src0] ^ XOR_j table row (j, bytes[j]); the nbytes bytes go where
 * other tasks keep their source list */
static void b_tab_ex(builder *b, uint32_t dst, uint32_t extra, int aux, uint32_t level, const uint8_t *bytes,
                     uint32_t nbytes);
static void b_tab(builder *b, uint32_t dst, uint32_t src0, uint32_t level, const uint8_t *bytes, uint32_t nbytes) {
  b_tab_ex(b, dst, src0, 0, level, bytes, nbytes);
}
/* extra: the row src0 (HBM flavour) / the arena row that also receives the result (smem flavour);
 * aux: bits per table group (smem flavour) */
static void b_tab_ex(builder *b, uint32_t dst, uint32_t src0, int aux, uint32_t level, const uint8_t *bytes,
                     uint32_t nbytes) {
  const uint32_t words = (nbytes + 3) / 4;
  b->tasks = sc_grow(b->sc, SC_TASKS, (b->nt + 1) * sizeof(ptask));
  b->srcs = sc_grow(b->sc, SC_SRCS, (b->ns + words + 1) * sizeof(uint32_t));
  if (b->sc->oom) return;
  ptask *t = &b->tasks[b->nt++];
  t->dst = dst;
  t->src_at = (uint32_t)b->ns;
  t->level = level;
  t->extra = src0;
  t->nsrc = (uint16_t)nbytes;
  t->kind = RQB_T_TAB;
  t->aux = (uint8_t)aux;
  b->srcs[b->ns + words - 1] = 0;
  memcpy(b->srcs + b->ns, bytes, nbytes);
  b->ns += words;
  if (level > b->max_level) b->max_level = level;
  for (uint32_t k = 0; k < nbytes; k++) b->tot_x += bytes[k] != 0;
  b->tot_x += 1;
}

/* smem flavour: slots dst .. dst+count-1 = arena rows first_row .. first_row+count-1 */
static void b_load(builder *b, uint32_t dst, uint32_t first_row, uint32_t count, uint32_t level) {
  b->tasks = sc_grow(b->sc, SC_TASKS, (b->nt + 1) * sizeof(ptask));
  if (b->sc->oom) return;
  ptask *t = &b->tasks[b->nt++];
  t->dst = dst;
  t->src_at = (uint32_t)b->ns;
  t->level = level;
  t->extra = first_row;
  t->nsrc = (uint16_t)count;
  t->kind = RQB_T_LOAD;
  t->aux = 0;
  if (level > b->max_level) b->max_level = level;
  b->tot_x += count;
}
/* smem flavour: one chunk of the alpha-scan with the HDPC sums folded in (rqb_program.h) */
static void b_scan2(builder *b, uint32_t acc_base, uint32_t yend, int aux, uint32_t level, const uint32_t *entries,
                    uint32_t n) {
  b_task(b, RQB_T_SCAN2, acc_base, aux, level, entries, n);
  if (!b->sc->oom) b->tasks[b->nt - 1].extra = yend;
}

/* row[dst] = sum of n sources that all exist before `level`: one task, or a
 * tree of partial sums into fresh working rows (leaves keep `kind`, inner nodes
 * are plain XORs).  srcs is clobbered.  Returns the level that writes dst. */
static uint32_t b_tree(builder *b, int kind, uint32_t dst, uint32_t *srcs, uint32_t n, uint32_t level) {
  if (n > RQB_MAX_SRCS && n <= RQB_MAX_SRCS * RQB_MAX_SRCS) {
    /* two levels with the fewest tasks, ceil((n-1)/7): k partial sums over c = n+k-8 of the
     * sources, then one task over the 8-k sources left and the k partial sums */
    const uint32_t k = (n - RQB_MAX_SRCS + RQB_MAX_SRCS - 2) / (RQB_MAX_SRCS - 1), c = n + k - RQB_MAX_SRCS;
    uint32_t part[RQB_MAX_SRCS], o = 0;
    for (uint32_t i = 0; i < k; i++) {
      uint32_t cnt = (c - o) / (k - i); /* spread evenly; every partial gets >= 2 sources */
      uint32_t r = b->ws_base + b->ws_next++;
      b_task(b, kind, r, 0, level, srcs + o, cnt);
      part[i] = RQB_SRC(r, 1);
      o += cnt;
    }
    uint32_t last[RQB_MAX_SRCS], m = 0;
    for (uint32_t q = c; q < n; q++) last[m++] = srcs[q];
    for (uint32_t i = 0; i < k; i++) last[m++] = part[i];
    b_task(b, kind, dst, 0, level + 1, last, m);
    return level + 1;
  }
  while (n > RQB_MAX_SRCS) {
    uint32_t m = 0;
    for (uint32_t o = 0; o < n; o += RQB_MAX_SRCS) {
      uint32_t cnt = n - o < RQB_MAX_SRCS ? n - o : RQB_MAX_SRCS;
      uint32_t r = b->ws_base + b->ws_next++;
      b_task(b, kind, r, 0, level, srcs + o, cnt);
      srcs[m++] = RQB_SRC(r, 1);
    }
    n = m;
    kind = RQB_T_XOR;
    level++;
  }
  b_task(b, kind, dst, 0, level, srcs, n);
  return level;
}

/* bytes of a task's source list in the page: XOR lists are padded to 4 or 8 entries */
static size_t list_bytes(const ptask *t) {
  if (t->kind == RQB_T_XOR) return t->nsrc <= 4 ? 16 : 32;
  if (t->kind == RQB_T_LOAD) return 0;
  if (t->kind == RQB_T_TAB) return ((size_t)t->nsrc + 15) & ~(size_t)15;
  return ((size_t)t->nsrc * 4 + 15) & ~(size_t)15;
}
static size_t task_bytes(const ptask *t) { return sizeof(rqb_task) + list_bytes(t); }

/* pack the tasks, level by level, into pages (a level may be split over pages:
 * its tasks are independent).  Returns 0 or a negative error. */
static int write_pages(builder *b, rqb_plan *plan, uint32_t zero_row, int smem, size_t *tot_levels, uint8_t *ext_buf,
                       size_t ext_cap) {
  scratch_t *sc = b->sc;
  const uint32_t nl = b->max_level + 1;
  /* one stable counting sort by (level, kind, narrow/wide list): tasks that take the
   * same path through the kernel sit together inside a level */
  const uint32_t nkeys = nl * 8;
  uint32_t *cnt = sc_buf(sc, SC_LVLCNT, ((size_t)nkeys + 2) * 4, 1);
  uint32_t *order = sc_buf(sc, SC_ORDER, (b->nt + 1) * 4, 0);
#define TKEY(t) ((t).level * 8 + ((t).kind == RQB_T_XOR ? (uint32_t)((t).nsrc > 4) : (t).kind + 1u))
  size_t tot_bytes = 0;
  for (size_t k = 0; k < b->nt; k++) {
    cnt[TKEY(b->tasks[k]) + 1]++;
    tot_bytes += task_bytes(&b->tasks[k]);
  }
  for (uint32_t l = 0; l < nkeys; l++) cnt[l + 1] += cnt[l];
  for (size_t k = 0; k < b->nt; k++) order[cnt[TKEY(b->tasks[k])]++] = (uint32_t)k;
#undef TKEY
  /* cnt[key] now holds the END of its bucket: level l spans [cnt[8l-1], cnt[8l+7]) */
  size_t npages = 0, cur = 0, levels_in_page = 0, levels = 0;
  /* into the caller's buffer when the program certainly fits: the tasks, a header per level
   * (levels split over pages get one per piece) and up to 512 unused bytes at the end of a page */
  const size_t worst = ((tot_bytes + (size_t)nl * 32) / (RQB_PAGE_BYTES - 512) + 2) * RQB_PAGE_BYTES;
  int ext = ext_buf && worst <= ext_cap;
  uint8_t *pages = ext ? ext_buf : plan->own_pages;
#define OPEN_PAGE()                                                         \
  do {                                                                      \
    if (ext && (npages + 1) * RQB_PAGE_BYTES > ext_cap) {                   \
      /* the estimate was too small for this mix of task sizes: carry on in the plan's own buffer */ \
      if (npages * RQB_PAGE_BYTES > plan->pages_cap) {                      \
        uint8_t *np_ = realloc(plan->own_pages, (npages + 64) * 2 * RQB_PAGE_BYTES); \
        if (!np_) return -7;                                                \
        plan->own_pages = np_;                                              \
        plan->pages_cap = (npages + 64) * 2 * RQB_PAGE_BYTES;               \
      }                                                                     \
      memcpy(plan->own_pages, pages, npages * RQB_PAGE_BYTES);              \
      pages = plan->own_pages;                                              \
      ext = 0;                                                              \
    }                                                                       \
    if (!ext && (npages + 1) * RQB_PAGE_BYTES > plan->pages_cap) {          \
      uint8_t *np_ = realloc(plan->own_pages, (npages + 64) * 2 * RQB_PAGE_BYTES); \
      if (!np_) return -7;                                                  \
      plan->pages_cap = (npages + 64) * 2 * RQB_PAGE_BYTES;                 \
      pages = plan->own_pages = np_;                                        \
    }                                                                       \
    memset(pages + npages * RQB_PAGE_BYTES, 0, RQB_PAGE_BYTES);             \
    npages++;                                                               \
    cur = sizeof(rqb_page_hdr);                                             \
    levels_in_page = 0;                                                     \
  } while (0)
#define CLOSE_PAGE()                                                                              \
  do {                                                                                            \
    if (cur) ((rqb_page_hdr *)(pages + (npages - 1) * RQB_PAGE_BYTES))->n_levels = (uint32_t)levels_in_page; \
    cur = 0;                                                                                      \
  } while (0)
  for (uint32_t l = 0; l < nl; l++) {
    size_t lo = l ? cnt[8 * l - 1] : 0, hi = cnt[8 * l + 7];
    if (lo == hi) continue;
    size_t idx = lo;
    while (idx < hi) {
      if (!cur) OPEN_PAGE();
      size_t avail = RQB_PAGE_BYTES - cur, need = sizeof(rqb_level_hdr), j = idx;
      while (j < hi) {
        size_t add = task_bytes(&b->tasks[order[j]]);
        if (need + add > avail) break;
        need += add;
        j++;
      }
      if (j == idx) {
        if (cur == sizeof(rqb_page_hdr)) return -5; /* a single task larger than a page */
        CLOSE_PAGE();
        continue;
      }
      uint8_t *page = pages + (npages - 1) * RQB_PAGE_BYTES;
      rqb_level_hdr *lh = (rqb_level_hdr *)(page + cur);
      size_t n = j - idx;
      lh->n_tasks = (uint32_t)n;
      lh->next_off = (uint32_t)(cur + need);
      lh->tab_base = b->tab_base;
      lh->zero_row = zero_row;
      rqb_task *dst = (rqb_task *)(page + cur + sizeof(rqb_level_hdr));
      uint32_t soff = (uint32_t)(cur + sizeof(rqb_level_hdr) + n * sizeof(rqb_task));
      for (size_t k = 0; k < n; k++) {
        const ptask *t = &b->tasks[order[idx + k]];
        size_t sb = list_bytes(t);
        if (t->kind != RQB_T_LOAD)
          memcpy(page + soff, b->srcs + t->src_at, t->kind == RQB_T_TAB ? (size_t)t->nsrc : (size_t)t->nsrc * 4);
        uint8_t aux = t->aux;
        if (t->kind == RQB_T_XOR) {
          for (size_t q = t->nsrc; q < sb / 4; q++) ((uint32_t *)(page + soff))[q] = zero_row;
          if (smem) { /* aux bit 0: destination and every source are slots (the kernel's lean path) */
            int all_slots = !(t->dst & RQB_REF_GLOBAL);
            for (size_t q = 0; q < t->nsrc; q++) all_slots &= !(b->srcs[t->src_at + q] & RQB_REF_GLOBAL);
            aux = (uint8_t)all_slots;
          }
        }
        dst[k].src_off = soff;
        dst[k].dst = t->dst;
        dst[k].nsrc = t->nsrc;
        dst[k].kind = t->kind;
        dst[k].aux = aux;
        dst[k].pad = t->extra;
        soff += (uint32_t)sb;
      }
      cur += need;
      levels_in_page++;
      levels++;
      idx = j;
      if (RQB_PAGE_BYTES - cur < sizeof(rqb_level_hdr) + sizeof(rqb_task) + 16) CLOSE_PAGE();
    }
  }
  CLOSE_PAGE();
#undef OPEN_PAGE
#undef CLOSE_PAGE
  plan->n_pages = (uint32_t)npages;
  plan->pages = pages;
  *tot_levels = levels;
  return 0;
}

/* ------------------------------------------------------------ bit helpers */
static inline void bits_xor(uint64_t *a, const uint64_t *b, int w) {
  switch (w) { /* rows of the bit matrices are a few words long: no loop for the common sizes */
    case 4: a[3] ^= b[3]; /* fall through */
    case 3: a[2] ^= b[2]; /* fall through */
    case 2: a[1] ^= b[1]; /* fall through */
    case 1: a[0] ^= b[0]; return;
    default:
      for (int k = 0; k < w; k++) a[k] ^= b[k];
  }
}
static inline int bit_get(const uint64_t *a, int k) { return (int)(a[k >> 6] >> (k & 63)) & 1; }
static inline void bit_flip(uint64_t *a, int k) { a[k >> 6] ^= (uint64_t)1 << (k & 63); }

/* packed-byte alpha multiply (x -> 2x in GF(256)) on 8 bytes at once */
static inline uint64_t xtime8(uint64_t x) {
  return ((x & 0x7f7f7f7f7f7f7f7fULL) << 1) ^ (((x >> 7) & 0x0101010101010101ULL) * 0x1d);
}

/* ------------------------------------------------------- plan object pool */
static rqb_plan *g_free_plans;
static pthread_mutex_t g_plans_mu = PTHREAD_MUTEX_INITIALIZER;

static rqb_plan *plan_acquire(void) {
  pthread_mutex_lock(&g_plans_mu);
  rqb_plan *p = g_free_plans;
  if (p) g_free_plans = p->next_free;
  pthread_mutex_unlock(&g_plans_mu);
  if (!p) p = calloc(1, sizeof(*p));
  return p;
}

/* really free the recycled plan objects and the per-K' matrix cache (rqb_release_cached) */
static void base_drain(void);
void rqb_plan_pool_drain(void) {
  pthread_mutex_lock(&g_plans_mu);
  rqb_plan *p = g_free_plans;
  g_free_plans = NULL;
  pthread_mutex_unlock(&g_plans_mu);
  while (p) {
    rqb_plan *n = p->next_free;
    free(p->own_pages);
    free(p);
    p = n;
  }
  base_drain();
}

void rqb_plan_free(rqb_plan *p) {
  if (!p) return;
  pthread_mutex_lock(&g_plans_mu); /* keeps its page buffer for the next block */
  p->next_free = g_free_plans;
  g_free_plans = p;
  pthread_mutex_unlock(&g_plans_mu);
}

/* ------------------------------------------------- per-K' matrix cache
 * The part of the constraint matrix that does not depend on which symbols were
 * received: the S LDPC rows (lib/precode.c:34-58) and the LT row of every
 * source/padding ISI 0..K'-1 (params_set_idxs, lib/params.c:47-65), as CSR, plus
 * each row's number of non-zeros in the columns [0, W).  Built once per K'. */
typedef struct base_matrix {
  int Kp;
  int *rptr; /* S + H + K' + 1 */
  int *cidx;
  int *deg;  /* S + H + K' */
  uint8_t *hb1, *hb2; /* K'+S each: the HDPC rows that get a one in column j (lib/precode.c:68-81) */
  struct base_matrix *next;
} base_matrix;
static base_matrix *g_base[64];
static pthread_mutex_t g_base_mu = PTHREAD_MUTEX_INITIALIZER;

static const base_matrix *base_get(const rqb_params *P) {
  const int S = P->S, H = P->H, W = P->W, Kp = P->Kprime, B = P->B;
  base_matrix **slot = &g_base[(unsigned)Kp % 64u];
  pthread_mutex_lock(&g_base_mu); /* held for a list walk of a few entries; building happens once per K' */
  base_matrix *m = *slot;
  while (m && m->Kp != Kp) m = m->next;
  if (!m) {
    const int rows = S + H + Kp;
    const int n = Kp + S;
    m = calloc(1, sizeof(*m));
    int *cur = malloc(sizeof(int) * (size_t)(S + 1));
    if (m) {
      m->Kp = Kp;
      m->rptr = calloc((size_t)rows + 2, sizeof(int));
      m->deg = calloc((size_t)rows + 1, sizeof(int));
      m->cidx = malloc(sizeof(int) * ((size_t)3 * B + 3 * (size_t)S + (size_t)RQB_MAX_LT_DEGREE * (size_t)Kp + 16));
      m->hb1 = calloc((size_t)n + 1, 1);
      m->hb2 = calloc((size_t)n + 1, 1);
    }
    if (!m || !cur || !m->rptr || !m->deg || !m->cidx || !m->hb1 || !m->hb2) { /* out of memory: nothing is cached */
      if (m) {
        free(m->rptr);
        free(m->deg);
        free(m->cidx);
        free(m->hb1);
        free(m->hb2);
      }
      free(m);
      free(cur);
      pthread_mutex_unlock(&g_base_mu);
      return NULL;
    }
    int *rptr = m->rptr, *cidx = m->cidx;
    for (int col = 0; col < B; col++) {
      int sub = col / S;
      rptr[1 + col % S]++;
      rptr[1 + (col + sub + 1) % S]++;
      rptr[1 + (col + 2 * (sub + 1)) % S]++;
    }
    for (int r = 0; r < S; r++) rptr[1 + r] += 3;
    int acc = 0;
    for (int r = 0; r < S; r++) {
      int c = rptr[1 + r];
      rptr[r] = acc;
      acc += c;
    }
    for (int r = S; r <= S + H; r++) rptr[r] = acc;
    memcpy(cur, rptr, sizeof(int) * (size_t)S);
    for (int col = 0; col < B; col++) {
      int sub = col / S;
      cidx[cur[col % S]++] = col;
      cidx[cur[(col + sub + 1) % S]++] = col;
      cidx[cur[(col + 2 * (sub + 1)) % S]++] = col;
    }
    for (int r = 0; r < S; r++) {
      cidx[cur[r]++] = B + r;
      cidx[cur[r]++] = W + r % P->P;
      cidx[cur[r]++] = W + (r + 1) % P->P;
    }
    free(cur);
    uint32_t idx[RQB_MAX_LT_DEGREE];
    for (int k = 0; k < Kp; k++) {
      int cnt = rqb_host_lt_indices(P, (uint32_t)k, idx);
      for (int q = 0; q < cnt; q++) cidx[acc + q] = (int)idx[q];
      acc += cnt;
      rptr[S + H + k + 1] = acc;
    }
    for (int r = 0; r < rows; r++)
      for (int k = rptr[r]; k < rptr[r + 1]; k++) m->deg[r] += (cidx[k] < W);
    for (int j = 0; j + 1 < n; j++) {
      uint32_t b1 = rqb_rand(rqb_rand_v, (uint32_t)j + 1, 6, (uint32_t)H);
      uint32_t b2 = (b1 + rqb_rand(rqb_rand_v, (uint32_t)j + 1, 7, (uint32_t)H - 1) + 1) % (uint32_t)H;
      m->hb1[j] = (uint8_t)b1;
      m->hb2[j] = (uint8_t)b2;
    }
    m->next = *slot;
    *slot = m;
  }
  pthread_mutex_unlock(&g_base_mu);
  return m;
}

static void base_drain(void) {
  pthread_mutex_lock(&g_base_mu);
  for (int k = 0; k < 64; k++) {
    base_matrix *m = g_base[k];
    g_base[k] = NULL;
    while (m) {
      base_matrix *n = m->next;
      free(m->rptr);
      free(m->cidx);
      free(m->deg);
      free(m->hb1);
      free(m->hb2);
      free(m);
      m = n;
    }
  }
  pthread_mutex_unlock(&g_base_mu);
}

/* sort the n pairs (rd[k], it[k]) by rd ascending, stable.  Short lists by insertion;
 * long ones (LDPC rows carry ~90 terms) by an 8-bit LSD radix over the level. */
static void sort_by_level(int *rd, int *it, int n, int maxkey, uint64_t *tmp /* 2n */) {
  if (n <= 40) {
    for (int a = 1; a < n; a++) {
      int lv = rd[a], q = it[a], j = a;
      while (j > 0 && rd[j - 1] > lv) {
        rd[j] = rd[j - 1];
        it[j] = it[j - 1];
        j--;
      }
      rd[j] = lv;
      it[j] = q;
    }
    return;
  }
  uint64_t *x = tmp, *y = tmp + n;
  for (int k = 0; k < n; k++) x[k] = ((uint64_t)(uint32_t)rd[k] << 32) | (uint32_t)it[k];
  for (int shift = 32; shift < 64 && (maxkey >> (shift - 32)) != 0; shift += 8) {
    int cnt[257] = {0};
    for (int k = 0; k < n; k++) cnt[((x[k] >> shift) & 255u) + 1]++;
    for (int k = 0; k < 256; k++) cnt[k + 1] += cnt[k];
    for (int k = 0; k < n; k++) y[cnt[(x[k] >> shift) & 255u]++] = x[k];
    uint64_t *t = x;
    x = y;
    y = t;
  }
  for (int k = 0; k < n; k++) {
    rd[k] = (int)(x[k] >> 32);
    it[k] = (int)(uint32_t)x[k];
  }
}

/* NANORQ_B200_BACKSUB = tables | triangular overrides the planner's choice of
 * back-substitution (experiments); read once */
static int g_fr_mode;
static pthread_once_t backsub_once = PTHREAD_ONCE_INIT;
static void backsub_init(void) {
  const char *e = getenv("NANORQ_B200_BACKSUB");
  g_fr_mode = !e ? 0 : !strncmp(e, "tables", 6) ? 1 : !strcmp(e, "triangular") ? 2 : 0;
}

/* ------------------------------------------------- shared-memory flavour
 * In-place chains: a row that lives in a shared-memory slot accumulates its terms where it
 * is, up to 7 at a time (8 sources per task, one of them the row itself), each group as early
 * as its terms exist -- no rows for partial sums.  rd[]/it[] = level at which each term is
 * final / the term; returns the level at which the row is final. */
typedef struct {
  int *first;  /* row index -> first group; rows 0..I-1 peeled positions, I..I+nb-1 residual rows */
  int *level;  /* group -> level of its task */
  int *ptr;    /* group -> first item */
  int *items;  /* peeled positions */
  int ng, nitems;
} chains;

static int chain_row(chains *ch, int *rd, int *it, int cnt, int has_self, int maxkey, uint64_t *sortbuf) {
  int prev = 0, idx = 0, cap = has_self ? (int)RQB_MAX_SRCS - 1 : (int)RQB_MAX_SRCS;
  if (cnt <= cap) { /* the common case: one task, only the latest term matters */
    if (cnt == 0) {
      ch->ptr[ch->ng] = ch->nitems;
      return 0;
    }
    int mx = rd[0];
    for (int k = 0; k < cnt; k++) {
      ch->items[ch->nitems + k] = it[k];
      if (rd[k] > mx) mx = rd[k];
    }
    ch->level[ch->ng] = mx + 1;
    ch->ptr[ch->ng] = ch->nitems;
    ch->nitems += cnt;
    ch->ng++;
    ch->ptr[ch->ng] = ch->nitems;
    return mx + 1;
  }
  sort_by_level(rd, it, cnt, maxkey, sortbuf);
  while (idx < cnt) {
    const int take = cnt - idx < cap ? cnt - idx : cap;
    const int lv = (rd[idx + take - 1] > prev ? rd[idx + take - 1] : prev) + 1;
    ch->level[ch->ng] = lv;
    ch->ptr[ch->ng] = ch->nitems;
    for (int k = 0; k < take; k++) ch->items[ch->nitems++] = it[idx + k];
    ch->ng++;
    prev = lv;
    idx += take;
    cap = (int)RQB_MAX_SRCS - 1;
  }
  ch->ptr[ch->ng] = ch->nitems;
  return prev;
}

/* bits [bits*j, bits*j + bits) of a bit row of uw words */
static inline uint32_t group_val(const uint64_t *g, int uw, int j, int bits) {
  const int off = j * bits, w = off >> 6, sh = off & 63;
  uint64_t x = g[w] >> sh;
  if (sh + bits > 64 && w + 1 < uw) x |= g[w + 1] << (64 - sh);
  return (uint32_t)(x & ((1u << bits) - 1u));
}

/* what the analysis (steps 1-3 of rqb_plan_build) hands to the emitter */
typedef struct {
  const rqb_plan_request *req;
  rqb_params P;
  int oh, R, n, nlt, I, U, uw, nb, nbw, uq, nfree, maxlevel;
  const int *rptr, *cidx, *col_pos, *col_t, *prow, *pcol, *ucol, *lowrows, *pivrow, *freecols, *qrow_of_f;
  const uint8_t *col_state, *hb1, *hb2, *Sh, *TQ;
  const uint64_t *G, *Sb, *Tb;
  const chains *ch;
  uint32_t slot_budget; /* slots a CTA can hold at the chosen slice width */
  uint32_t row0[4];
} plan_view;

#define SM_G(row) (RQB_REF_GLOBAL | (uint32_t)(row))

/* Emits the shared-memory flavour of the program (rqb_program.h).  Slots:
 *   0                 zeros
 *   1 + r             matrix row r (LDPC, HDPC, LT), updated in place: Y_p, r_m, r'_h, finally x_p
 *   1 + R + t         inactive column t: r'_t, then z_t in place
 *   1 + R + U ...     scratch, reused phase by phase: scan accumulators and y ends, partial sums
 *                     of the GF(256) trees, then the XOR tables of the back-substitution, then the
 *                     partial sums of long output rows
 * Returns 0, -7 (out of memory) or -8 (does not fit the slot budget: use the HBM flavour). */
static int emit_smem(const plan_view *v, builder *bd, scratch_t *sc, uint32_t *n_slots_out, uint32_t *tab_bits_out) {
#ifdef RQB_PLAN_FINE
  double fine_t = now_s();
#endif
  const rqb_plan_request *req = v->req;
  const int S = v->P.S, H = v->P.H, L = v->P.L, R = v->R, U = v->U, I = v->I, n = v->n, nb = v->nb;
  const int uw = v->uw, nbw = v->nbw, nfree = v->nfree;
  const uint32_t SCR = 1u + (uint32_t)R + (uint32_t)U;
  if (v->slot_budget <= SCR + 64) return -8;
  const uint32_t scr_cap = v->slot_budget - SCR;
#define SLOT_ROW(r) (1u + (uint32_t)(r))
#define SLOT_Z(t) (1u + (uint32_t)R + (uint32_t)(t))
  uint8_t *has = sc_buf(sc, SC_HAS, (size_t)SCR + 8, 1); /* the slot holds a value (otherwise: known zero) */
  uint32_t *tmp = sc_buf(sc, SC_TMP, sizeof(uint32_t) * ((size_t)L + (size_t)nb + (size_t)4 * (size_t)n + 4096), 0);
  uint8_t *needed = sc_buf(sc, SC_NEEDED, (size_t)L + 8, 0); /* columns whose intermediate symbol is an output term */
  if (sc->oom) return -7;
  uint32_t peak = SCR, sp;
  bd->ws_base = 0; /* b_tree takes rows for partial sums from ws_base + ws_next++: scratch slots here */
#define PUSHS(ns, slot)                                   \
  do {                                                    \
    uint32_t _s = (slot);                                 \
    if (has[_s]) tmp[(ns)++] = RQB_SRC(_s, 1);            \
  } while (0)

  /* which intermediate symbols are wanted: all of them (encoder), or the terms of the outputs */
  if (req->want_c) {
    memset(needed, 1, (size_t)L);
  } else {
    memset(needed, 0, (size_t)L);
    for (int k = 0; k < req->n_out; k++) {
      uint32_t idx[RQB_MAX_LT_DEGREE];
      int cnt = rqb_host_lt_indices(&v->P, req->out_isi[k], idx);
      for (int q = 0; q < cnt; q++) needed[idx[q]] = 1;
    }
  }

  /* level 0: the received symbols, HBM -> slots, in runs of consecutive rows */
  for (int k = 0; k < v->nlt;) {
    if (req->in_row[k] == RQB_ROW_NONE) {
      k++;
      continue;
    }
    if (req->in_row[k] >= req->in_rows) return -1;
    int len = 1;
    while (k + len < v->nlt && len < 16 && req->in_row[k + len] == req->in_row[k] + (uint32_t)len) len++;
    if (req->in_row[k] + (uint32_t)len > req->in_rows) return -1;
    b_load(bd, SLOT_ROW(S + H + k), SM_G(v->row0[RQB_SP_IN] + req->in_row[k]), (uint32_t)len, 0);
    for (int q = 0; q < len; q++) has[SLOT_ROW(S + H + k + q)] = 1;
    k += len;
  }

  FINE(15);
  /* A + B: the triangular solve Y = X^-1 b_top and the residual rows r_m = b_m ^ X_low Y, all in place */
  uint32_t endA = (uint32_t)v->maxlevel;
  for (int idx = 0; idx < I + nb; idx++) {
    const int r = idx < I ? v->prow[idx] : v->lowrows[idx - I];
    const uint32_t dst = SLOT_ROW(r);
    for (int g = v->ch->first[idx]; g < v->ch->first[idx + 1]; g++) {
      uint32_t ns = 0;
      PUSHS(ns, dst);
      const uint32_t self = ns;
      for (int e = v->ch->ptr[g]; e < v->ch->ptr[g + 1]; e++) PUSHS(ns, SLOT_ROW(v->prow[v->ch->items[e]]));
      if (ns == self) continue; /* nothing to add */
      b_task(bd, RQB_T_XOR, dst, 0, (uint32_t)v->ch->level[g], tmp, ns);
      has[dst] = 1;
      if ((uint32_t)v->ch->level[g] > endA) endA = (uint32_t)v->ch->level[g];
    }
  }
  const uint32_t lvB = (uint32_t)v->maxlevel + 1;
  const uint32_t endB = endA > lvB ? endA : lvB;
  FINE(16);

  /* alpha-scans over the columns 0..n-1 in NC chunks, HDPC sums accumulated on the way (SCAN2) */
  int NC = n / 32;
  if (NC > 128) NC = 128;
  if (NC < 1) NC = 1;
  for (;;) {
    const uint32_t gf_tmp = (uint32_t)H * (((uint32_t)NC + (uint32_t)U + 2u + 6u) / 7u + 3u);
    if ((uint32_t)NC * ((uint32_t)H + 1u) + gf_tmp <= scr_cap || NC == 1) break;
    NC /= 2;
  }
  if ((uint32_t)NC * ((uint32_t)H + 1u) + (uint32_t)H * 8u > scr_cap) return -8;
#define ACC(c, h) (SCR + (uint32_t)(c) * (uint32_t)H + (uint32_t)(h))
#define YEND(c) (SCR + (uint32_t)NC * (uint32_t)H + (uint32_t)(c))
  int *cs = sc_buf(sc, SC_CS, sizeof(int) * ((size_t)NC + 2), 0);
  if (sc->oom) return -7;
  for (int c = 0; c <= NC; c++) cs[c] = (int)((long)n * c / NC);
  for (int c = 0; c < NC; c++) {
    uint32_t ns = 0;
    for (int j = cs[c]; j < cs[c + 1]; j++) {
      uint32_t ref = RQB_REF_NONE;
      if (v->col_state[j] == 1 && has[SLOT_ROW(v->prow[v->col_pos[j]])]) ref = SLOT_ROW(v->prow[v->col_pos[j]]);
      const uint32_t h1 = j + 1 < n ? v->hb1[j] : 0u, h2 = j + 1 < n ? v->hb2[j] : 0u;
      tmp[ns++] = ref | (h1 << 24) | (h2 << 28);
    }
    b_scan2(bd, ACC(c, 0), YEND(c), H | (cs[c + 1] == n ? 0x80 : 0), lvB, tmp, ns);
  }
  sp = SCR + (uint32_t)NC * ((uint32_t)H + 1u);
  if (sp > peak) peak = sp;

  /* C1: r'_t = XOR of the residual rows the GF(2) elimination combined for pivot column t */
  uint32_t lvC = endB + 1, endC1 = endB;
  for (int t = 0; t < U; t++) {
    if (v->pivrow[t] < 0) continue;
    const uint64_t *tb = v->Tb + (size_t)v->pivrow[t] * nbw;
    const uint32_t dst = SLOT_Z(t);
    uint32_t ns = 0, lv = lvC;
    for (int m = 0; m < nb; m++) {
      if (!bit_get(tb, m) || !has[SLOT_ROW(v->lowrows[m])]) continue;
      tmp[ns++] = RQB_SRC(SLOT_ROW(v->lowrows[m]), 1);
      if (ns == RQB_MAX_SRCS) {
        b_task(bd, RQB_T_XOR, dst, 0, lv, tmp, ns);
        has[dst] = 1;
        if (lv > endC1) endC1 = lv;
        lv++;
        ns = 0;
        tmp[ns++] = RQB_SRC(dst, 1);
      }
    }
    if (ns > (has[dst] ? 1u : 0u)) {
      b_task(bd, RQB_T_XOR, dst, 0, lv, tmp, ns);
      has[dst] = 1;
      if (lv > endC1) endC1 = lv;
    }
  }
  /* HDPC sums: XOR the chunks' accumulators together, in place, 8 at a time */
  uint32_t endT = endB;
  {
    uint32_t lv = lvC;
    for (int stride = 1; stride < NC; stride *= 8, lv++) {
      for (int c0 = 0; c0 < NC; c0 += stride * 8)
        for (int h = 0; h < H; h++) {
          uint32_t ns = 0;
          for (int k = 0; k < 8 && c0 + k * stride < NC; k++) tmp[ns++] = RQB_SRC(ACC(c0 + k * stride, h), 1);
          if (ns > 1) b_task(bd, RQB_T_XOR, ACC(c0, h), 0, lv, tmp, ns);
        }
      endT = lv;
    }
  }
  uint32_t lv = (endC1 > endT ? endC1 : endT) + 1, end = lv;

  /* C2: r'_h = r_h ^ sum_c Gc[h][c]*yend_c ^ sum_t beta[h][t]*r'_t (GF leaves, XOR tree); the
   * coefficients as in the HBM flavour (see there) */
  {
    uint8_t *coef = sc_buf(sc, SC_COEF, (size_t)NC * (size_t)H, 1);
    uint8_t *gc = sc_buf(sc, SC_GC, (size_t)NC * (size_t)H + 64, 0);
    if (sc->oom) return -7;
    for (int c = 0; c < NC; c++)
      for (int j = cs[c]; j < cs[c + 1] && j + 1 < n; j++) {
        uint8_t a = rqb_gf_pow2(&GF, j - cs[c] + 1);
        coef[c * H + v->hb1[j]] ^= a;
        coef[c * H + v->hb2[j]] ^= a;
      }
    uint8_t suf[RQB_MAX_H];
    memset(suf, 0, sizeof(suf));
    for (int c1 = NC - 1; c1 >= 0; c1--) {
      if (c1 + 1 < NC) {
        uint8_t a = rqb_gf_pow2(&GF, cs[c1 + 2] - cs[c1 + 1]);
        for (int h = 0; h < H; h++) suf[h] = rqb_gf_mul(&GF, suf[h], a) ^ coef[(c1 + 1) * H + h];
      }
      for (int h = 0; h < H; h++)
        gc[c1 * H + h] = rqb_gf_mul(&GF, rqb_gf_pow2(&GF, h), rqb_gf_pow2(&GF, n - cs[c1 + 1])) ^ suf[h];
    }
    bd->ws_next = sp;
    for (int h = 0; h < H; h++) {
      const uint8_t *row = v->Sh + (size_t)h * (size_t)v->uq * 8;
      uint32_t ns = 0;
      for (int c1 = 0; c1 < NC; c1++)
        if (gc[c1 * H + h]) tmp[ns++] = RQB_SRC(YEND(c1), gc[c1 * H + h]);
      for (int t = 0; t < U; t++)
        if (v->pivrow[t] >= 0 && row[t] && has[SLOT_Z(t)]) tmp[ns++] = RQB_SRC(SLOT_Z(t), row[t]);
      tmp[ns++] = RQB_SRC(ACC(0, h), 1);
      uint32_t e2 = b_tree(bd, RQB_T_GF, SLOT_ROW(S + h), tmp, ns, lv);
      has[SLOT_ROW(S + h)] = 1;
      if (e2 > end) end = e2;
    }
    if (bd->ws_next > peak) peak = bd->ws_next;
    if (peak > v->slot_budget) return -8;
  }
  lv = end + 1;
  end = lv;
  /* C3: z_f = sum_h TQ[qrow(f)][h] * r'_h */
  for (int f = 0; f < nfree; f++) {
    uint32_t ns = 0;
    for (int h = 0; h < H; h++) {
      uint8_t b = v->TQ[v->qrow_of_f[f] * H + h];
      if (b) tmp[ns++] = RQB_SRC(SLOT_ROW(S + h), b);
    }
    uint32_t e2 = b_tree(bd, RQB_T_GF, SLOT_Z(v->freecols[f]), tmp, ns, lv);
    has[SLOT_Z(v->freecols[f])] = 1;
    if (e2 > end) end = e2;
  }
  if (bd->ws_next > peak) peak = bd->ws_next;
  if (peak > v->slot_budget) return -8;
  lv = end + 1;
  end = lv;
  /* C4: z_t = r'_t ^ XOR_{f: bit} z_f for the pivot columns, in place */
  for (int t = 0; t < U; t++) {
    if (v->pivrow[t] < 0) continue;
    const uint64_t *ps = v->Sb + (size_t)v->pivrow[t] * uw;
    const uint32_t dst = SLOT_Z(t);
    uint32_t ns = 0, l2 = lv;
    PUSHS(ns, dst);
    for (int f = 0; f < nfree; f++) {
      if (!bit_get(ps, v->freecols[f])) continue;
      tmp[ns++] = RQB_SRC(SLOT_Z(v->freecols[f]), 1);
      if (ns == RQB_MAX_SRCS) {
        b_task(bd, RQB_T_XOR, dst, 0, l2, tmp, ns);
        has[dst] = 1;
        if (l2 > end) end = l2;
        l2++;
        ns = 0;
        tmp[ns++] = RQB_SRC(dst, 1);
      }
    }
    if (ns > (has[dst] ? 1u : 0u)) {
      b_task(bd, RQB_T_XOR, dst, 0, l2, tmp, ns);
      has[dst] = 1;
      if (l2 > end) end = l2;
    }
  }
  lv = end + 1;

  FINE(17);
  /* F: x_p = Y_p ^ (G z)_p through XOR tables over groups of `bits` inactive symbols, the largest
   * group size whose tables fit the scratch slots (everything that lived there is dead by now).
   * Level lv: entries made of the z themselves (one half of the group's bits); lv+1: entries
   * that combine a low and a high half; lv+2: one in-place TAB task per wanted row. */
  int bits = 8;
  while (bits > 4 && (((uint32_t)(U + bits - 1) / (uint32_t)bits) << bits) > scr_cap) bits--;
  const int ngr = (U + bits - 1) / bits;
  if (((uint32_t)ngr << bits) > scr_cap) return -8;
  const uint32_t TAB = SCR;
  bd->tab_base = TAB;
  if (TAB + ((uint32_t)ngr << bits) > peak) peak = TAB + ((uint32_t)ngr << bits);
  const int lo_bits = bits / 2;
  const uint32_t lo_mask = (1u << lo_bits) - 1u, tsize = 1u << bits;
  uint8_t *used = sc_buf(sc, SC_FRUSED, (size_t)ngr * tsize + 64, 1);
  uint8_t *gv = sc_buf(sc, SC_FRTAB, (size_t)ngr + 64, 0);
  if (sc->oom) return -7;
#define GROUP_VAL(g, j) group_val((g), uw, (j), bits)
  for (int p = 0; p < I; p++) {
    if (!needed[v->pcol[p]]) continue;
    const uint64_t *g = v->G + (size_t)p * uw;
    for (int j = 0; j < ngr; j++) used[(size_t)j * tsize + GROUP_VAL(g, j)] = 1;
  }
  for (int j = 0; j < ngr; j++) {
    uint8_t *use = used + (size_t)j * tsize;
    for (uint32_t m = 1; m < tsize; m++)
      if (use[m] == 1 && (m & lo_mask) && (m & ~lo_mask)) {
        if (!use[m & lo_mask]) use[m & lo_mask] = 2;
        if (!use[m & ~lo_mask]) use[m & ~lo_mask] = 2;
      }
    for (uint32_t m = 1; m < tsize; m++) {
      if (!use[m]) continue;
      const uint32_t slot = TAB + ((uint32_t)j << bits) + m;
      if ((m & lo_mask) && (m & ~lo_mask)) {
        uint32_t pair[2] = {TAB + ((uint32_t)j << bits) + (m & lo_mask), TAB + ((uint32_t)j << bits) + (m & ~lo_mask)};
        b_task(bd, RQB_T_XOR, slot, 0, lv + 1, pair, 2);
      } else { /* up to 4 of the z themselves (a z that is known zero is left out; none left: a zero row) */
        uint32_t ns = 0;
        for (int bit = 0; bit < bits; bit++)
          if (m >> bit & 1) {
            int t = bits * j + bit;
            if (t < U) PUSHS(ns, SLOT_Z(t));
          }
        b_task(bd, RQB_T_XOR, slot, 0, lv, tmp, ns);
      }
    }
  }
  for (int p = 0; p < I; p++) {
    const int col = v->pcol[p];
    if (!needed[col]) continue;
    const uint64_t *g = v->G + (size_t)p * uw;
    const uint32_t slot = SLOT_ROW(v->prow[p]);
    const uint32_t crow = req->want_c ? SM_G(v->row0[RQB_SP_C] + (uint32_t)col) : RQB_ROW_NONE;
    int any = 0;
    for (int j = 0; j < ngr; j++) {
      gv[j] = (uint8_t)GROUP_VAL(g, j);
      any |= gv[j];
    }
    if (!any) { /* x_p = Y_p */
      if (req->want_c) {
        uint32_t ns = 0;
        PUSHS(ns, slot);
        b_task(bd, RQB_T_XOR, crow, 0, lv + 2, tmp, ns);
      }
      continue;
    }
    if (!has[slot]) { /* Y_p is known zero: give the in-place gather a zero row to start from */
      b_task(bd, RQB_T_XOR, slot, 0, lv, tmp, 0);
      has[slot] = 1;
    }
    b_tab_ex(bd, slot, crow, bits, lv + 2, gv, (uint32_t)ngr);
  }
#undef GROUP_VAL
  FINE(18);
  if (req->want_c)
    for (int t = 0; t < U; t++) { /* the inactive symbols themselves */
      uint32_t ns = 0;
      PUSHS(ns, SLOT_Z(t));
      b_task(bd, RQB_T_XOR, SM_G(v->row0[RQB_SP_C] + (uint32_t)v->ucol[t]), 0, lv + 2, tmp, ns);
    }
  lv += 3;

  /* O: emitted symbols = LT combinations of the intermediate symbols, slots -> HBM.  Rows with
   * more than 8 terms take scratch slots for partial sums; when those run out the next rows
   * start two levels later and reuse them. */
  bd->ws_next = SCR;
  for (int k = 0; k < req->n_out; k++) {
    uint32_t idx[RQB_MAX_LT_DEGREE];
    int cnt = rqb_host_lt_indices(&v->P, req->out_isi[k], idx);
    uint32_t ns = 0;
    for (int q = 0; q < cnt; q++) {
      const int col = (int)idx[q];
      PUSHS(ns, v->col_state[col] == 1 ? SLOT_ROW(v->prow[v->col_pos[col]]) : SLOT_Z(v->col_t[col]));
    }
    if (ns > RQB_MAX_SRCS && bd->ws_next + 8u > v->slot_budget) {
      lv += 2;
      bd->ws_next = SCR;
    }
    b_tree(bd, RQB_T_XOR, SM_G(v->row0[RQB_SP_SYM] + (req->out_row ? req->out_row[k] : (uint32_t)k)), tmp, ns, lv);
    if (bd->ws_next > peak) peak = bd->ws_next;
  }
  FINE(19);
  if (peak > v->slot_budget) return -8;
  *n_slots_out = peak;
  *tab_bits_out = (uint32_t)bits;
  return sc->oom ? -7 : 0;
#undef SLOT_ROW
#undef SLOT_Z
#undef ACC
#undef YEND
#undef PUSHS
}

/* NANORQ_B200_SMEM_MAXSLICE = 16 | 32 | 64: widest column slice of the shared-memory flavour
 * (experiments); read once */
static uint32_t g_smem_max_slice = 64;
static pthread_once_t smem_once = PTHREAD_ONCE_INIT;
static void smem_init(void) {
  const char *e = getenv("NANORQ_B200_SMEM_MAXSLICE");
  const int v = e ? atoi(e) : 0;
  if (v == 16 || v == 32 || v == 64) g_smem_max_slice = (uint32_t)v;
}
/* arena rows below the working rows: [IN | SYM | C | ZERO] */
static uint64_t row0_ws_rows(const rqb_plan_request *req, int L) {
  return (uint64_t)req->in_rows + req->sym_rows + (uint64_t)L + 1u;
}

int (*rqb_plan_usolve_hook)(rqb_usolve_io *io);
int rqb_plan_usolve_mode;

/* ---------------------------------------------------------------- planner */
#define NONE_REF RQB_REF_NONE /* "this row is all zero / has no location" */

int rqb_plan_build(const rqb_plan_request *req, rqb_plan **out) {
  pthread_once(&tables_once, tables_build);
  *out = NULL;
  rqb_params P;
  if (rqb_params_init(req->K, &P) || req->overhead < 0) return -1;
  const int S = P.S, H = P.H, W = P.W, L = P.L, Kp = P.Kprime, B = P.B;
  const int oh = req->overhead, R = L + oh, n = Kp + S, nlt = Kp + oh;
  if ((uint32_t)H > RQB_MAX_H) return -1;
  double t0 = now_s();
#ifdef RQB_PLAN_FINE
  double fine_t = t0;
#endif
  scratch_t *sc = sc_get();
  int rc = 0;
  if (!sc) return -7;
  sc->oom = 0;
#define OOM_CHECK() do { if (sc->oom) return -7; } while (0)

  /* ---- 1. sparse matrix A, rows: [0,S) LDPC, [S,S+H) HDPC (kept empty, closed form),
   *         [S+H, R) LT rows.  Same contents as precode_matrix_gen (+patching). */
  const base_matrix *bm = base_get(&P);
  if (!bm) return -6;
  int *rptr = sc_buf(sc, SC_RPTR, ((size_t)R + 2) * sizeof(int), 0);
  size_t cap = (size_t)3 * B + 3 * (size_t)S + (size_t)RQB_MAX_LT_DEGREE * (size_t)nlt + 16;
  int *cidx = sc_buf(sc, SC_CIDX, cap * sizeof(int), 0);
  int *deg = sc_buf(sc, SC_DEG, (size_t)R * sizeof(int), 0); /* non-zeros in the columns [0, W) */
  OOM_CHECK();
  {
    /* LDPC rows and the LT rows of the source symbols (ISI k in row k) are the same for
     * every block of this K': copied from the per-K' cache; only the rows of repair
     * symbols are generated (Tuple + index walk, ~100 ns each) */
    const int fixed = bm->rptr[S + H];
    memcpy(rptr, bm->rptr, sizeof(int) * (size_t)(S + H + 1));
    memcpy(cidx, bm->cidx, sizeof(int) * (size_t)fixed);
    memcpy(deg, bm->deg, sizeof(int) * (size_t)(S + H));
    int acc = fixed;
    uint32_t idx[RQB_MAX_LT_DEGREE];
    for (int k = 0; k < nlt; k++) {
      const uint32_t x = req->isi[k];
      if (x < (uint32_t)Kp) {
        const int *src = bm->cidx + bm->rptr[S + H + x];
        const int cnt = bm->rptr[S + H + x + 1] - bm->rptr[S + H + x];
        for (int q = 0; q < cnt; q++) cidx[acc + q] = src[q];
        acc += cnt;
        deg[S + H + k] = bm->deg[S + H + x];
      } else {
        int cnt = rqb_host_lt_indices(&P, x, idx), dg = 0;
        for (int q = 0; q < cnt; q++) {
          cidx[acc + q] = (int)idx[q];
          dg += idx[q] < (uint32_t)W;
        }
        acc += cnt;
        deg[S + H + k] = dg;
      }
      rptr[S + H + k + 1] = acc;
    }
  }
  const int nnz = rptr[R];
  FINE(20);
  /* column lists */
  int *cptr = sc_buf(sc, SC_CPTR, ((size_t)L + 1) * sizeof(int), 1);
  int *ridx = sc_buf(sc, SC_RIDX, sizeof(int) * (size_t)(nnz ? nnz : 1), 0);
  int *cur_scratch = sc_buf(sc, SC_CUR, sizeof(int) * (size_t)(S > L ? S : L), 0);
  OOM_CHECK();
  for (int k = 0; k < nnz; k++) cptr[cidx[k] + 1]++;
  for (int c = 0; c < L; c++) cptr[c + 1] += cptr[c];
  {
    int *cur = cur_scratch;
    memcpy(cur, cptr, sizeof(int) * (size_t)L);
    for (int r = 0; r < R; r++)
      for (int k = rptr[r]; k < rptr[r + 1]; k++) ridx[cur[cidx[k]]++] = r;
  }
  double t1 = now_s();
  FINE(21);

  /* ---- 2. peeling */
  uint8_t *col_state = sc_buf(sc, SC_COLSTATE, (size_t)L, 1); /* 0 active, 1 peeled, 2 inactive */
  int *col_pos = sc_buf(sc, SC_COLPOS, sizeof(int) * (size_t)L, 0);
  int *col_t = sc_buf(sc, SC_COLT, sizeof(int) * (size_t)L, 0);
  int *row_pos = sc_buf(sc, SC_ROWPOS, sizeof(int) * (size_t)R, 0);
  int *prow = sc_buf(sc, SC_PROW, sizeof(int) * (size_t)L, 0);
  int *pcol = sc_buf(sc, SC_PCOL, sizeof(int) * (size_t)L, 0);
  int *ucol = sc_buf(sc, SC_UCOL, sizeof(int) * (size_t)L, 0);
  int *stk1 = sc_buf(sc, SC_STK1, sizeof(int) * ((size_t)nnz + (size_t)R + 8), 0);
  int *stk2 = sc_buf(sc, SC_STK2, sizeof(int) * ((size_t)nnz + (size_t)R + 8), 0);
  int n1 = 0, n2 = 0, ni = 0, nu = 0, h1 = 0, h2 = 0;
  OOM_CHECK();
  for (int r = 0; r < R; r++) row_pos[r] = -1;
  for (int c = 0; c < L; c++) col_pos[c] = col_t[c] = -1;
  for (int c = W; c < L; c++) { /* the P permanently inactive columns */
    col_state[c] = 2;
    col_t[c] = nu;
    ucol[nu++] = c;
  }
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
    /* rows with one column left are taken in the order they got there (a queue, not the
     * reference's LIFO bucket, lib/precode.c:115-126): breadth first, a row tends to depend on
     * pivots found early, and the dependency depth of the triangular solve -- the number of
     * levels the kernel needs -- drops from ~500 to ~250-350 at K=4096.  Same C either way. */
    while (h1 < n1) {
      int cand = stk1[h1++];
      if (row_pos[cand] < 0 && deg[cand] == 1) {
        r = cand;
        break;
      }
    }
    if (r < 0)
      while (h2 < n2) {
        int cand = stk2[h2++];
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
  if (I + U != L || nb < 0) return -2;
  double t2 = now_s();
  FINE(5);

  /* Flavour: when a CTA's shared memory can hold a column slice of every live row (the matrix
   * rows, the inactive symbols and a scratch region for accumulators / XOR tables), the
   * shared-memory program is emitted (rqb_program.h): the widest slice of 64, 32 or 16 bytes
   * that fits.  Larger blocks get the HBM flavour. */
  int smem = 0;
  uint32_t slice = 0, slot_budget = 0;
  if (req->smem_budget) {
    pthread_once(&smem_once, smem_init);
    const uint32_t fixed = 1u + (uint32_t)R + (uint32_t)U;
    uint32_t min_scratch = 16u * (((uint32_t)U + 3u) / 4u);                 /* 4-bit tables */
    const uint32_t scan_min = 8u * ((uint32_t)H + 1u) + (uint32_t)H * ((8u + (uint32_t)U + 8u) / 7u + 3u);
    if (scan_min > min_scratch) min_scratch = scan_min;
    if (min_scratch < 128u) min_scratch = 128u;
    for (uint32_t w = g_smem_max_slice; w >= 16u; w >>= 1)
      if ((uint64_t)(fixed + min_scratch) * w <= req->smem_budget && (uint64_t)row0_ws_rows(req, L) < RQB_REF_GLOBAL) {
        smem = 1;
        slice = w;
        slot_budget = req->smem_budget / w;
        break;
      }
  }

  /* ---- 3a. dependencies of the triangular solve, G = X^-1 U_top (bits), and the
   * schedule of each row: at most RQB_MAX_SRCS-1 terms besides the row itself;
   * longer rows get partial sums ("parts") placed at the earliest level their
   * terms exist.  Items: q >= 0 = peeled position q, -1-k = part k. */
  const int CAPF = (int)RQB_MAX_SRCS - 1;
  int *level = sc_buf(sc, SC_LEVEL, sizeof(int) * ((size_t)I + 1), 0);
  uint64_t *G = sc_buf(sc, SC_G, (size_t)(I ? I : 1) * uw * sizeof(uint64_t), 1);
  int *fptr = sc_buf(sc, SC_FPTR, sizeof(int) * ((size_t)I + 1), 0);
  int *fitems = sc_buf(sc, SC_FITEMS, sizeof(int) * ((size_t)I * CAPF + 8), 0);
  int *pfirst = sc_buf(sc, SC_PFIRST, sizeof(int) * ((size_t)I + 1), 0);
  /* parts: at most nnz/2 of them, items: at most nnz + #parts */
  int *plevel = sc_buf(sc, SC_PLEVEL, sizeof(int) * ((size_t)nnz / 2 + 8), 0);
  int *pptr = sc_buf(sc, SC_PPTR, sizeof(int) * ((size_t)nnz / 2 + 9), 0);
  int *pitems = sc_buf(sc, SC_PITEMS, sizeof(int) * ((size_t)nnz * 2 + 64), 0);
  int nparts = 0, npitems = 0, maxlevel = 0;
  /* one word per column instead of three arrays: >= 0 peeled at that position, < 0 inactive
   * with index ~value */
  int *cinfo = sc_buf(sc, SC_CINFO, sizeof(int) * (size_t)L, 0);
  chains ch;
  memset(&ch, 0, sizeof(ch));
  if (smem) { /* in-place chains instead of parts: at most one group per term */
    ch.first = sc_buf(sc, SC_CHFIRST, sizeof(int) * ((size_t)R + 2), 0);
    ch.level = sc_buf(sc, SC_CHLEVEL, sizeof(int) * ((size_t)nnz + (size_t)R + 8), 0);
    ch.ptr = sc_buf(sc, SC_CHPTR, sizeof(int) * ((size_t)nnz + (size_t)R + 9), 0);
    ch.items = sc_buf(sc, SC_CHITEMS, sizeof(int) * ((size_t)nnz + 8), 0);
  }
  OOM_CHECK();
  for (int c = 0; c < L; c++) cinfo[c] = col_state[c] == 1 ? col_pos[c] : ~col_t[c];
  {
    int nf = 0;
    int *it = sc_buf(sc, SC_TMP, sizeof(int) * 2 * ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64), 0);
    int *rd = it + ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64);
    uint64_t *sortbuf = sc_buf(sc, SC_SORT, sizeof(uint64_t) * 2 * ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64), 0);
    OOM_CHECK();
    pptr[0] = 0;
    for (int p = 0; p < I; p++) {
      int r = prow[p], cnt = 0;
      uint64_t *g = G + (size_t)p * uw;
      pfirst[p] = nparts;
      for (int k = rptr[r]; k < rptr[r + 1]; k++) {
        const int q = cinfo[cidx[k]];
        if (q < 0) {
          bit_flip(g, ~q);
        } else if (q != p) {
          if (q > p) return -3; /* cannot happen: would contradict the peeling invariant */
          bits_xor(g, G + (size_t)q * uw, uw);
          rd[cnt] = level[q];
          it[cnt++] = q;
        }
      }
      if (smem) {
        ch.first[p] = ch.ng;
        level[p] = chain_row(&ch, rd, it, cnt, r >= S + H && req->in_row[r - S - H] != RQB_ROW_NONE, maxlevel, sortbuf);
        if (level[p] > maxlevel) maxlevel = level[p];
        continue;
      }
      if (cnt <= CAPF) { /* the common case: one task, only the latest term matters */
        int mx = -1;
        fptr[p] = nf;
        for (int k = 0; k < cnt; k++) {
          fitems[nf++] = it[k];
          if (rd[k] > mx) mx = rd[k];
        }
        level[p] = mx + 1;
        if (level[p] > maxlevel) maxlevel = level[p];
        continue;
      }
      sort_by_level(rd, it, cnt, maxlevel, sortbuf); /* by the level at which the term exists */
      int head = 0; /* items [head, cnt) are live, sorted by readiness */
      while (cnt - head > CAPF) {
        int take = cnt - head - CAPF + 1;
        if (take > (int)RQB_MAX_SRCS) take = (int)RQB_MAX_SRCS;
        int lv = rd[head + take - 1] + 1;
        for (int k = 0; k < take; k++) pitems[npitems++] = it[head + k];
        plevel[nparts] = lv;
        pptr[nparts + 1] = npitems;
        head += take;
        /* put the part back, keeping the order */
        int j = head - 1;
        while (j + 1 < cnt && rd[j + 1] <= lv) {
          rd[j] = rd[j + 1];
          it[j] = it[j + 1];
          j++;
        }
        rd[j] = lv;
        it[j] = -1 - nparts;
        head--;
        nparts++;
      }
      fptr[p] = nf;
      for (int k = head; k < cnt; k++) fitems[nf++] = it[k];
      level[p] = cnt > head ? rd[cnt - 1] + 1 : 0;
      if (level[p] > maxlevel) maxlevel = level[p];
    }
    fptr[I] = nf;
    pfirst[I] = nparts;
  }

  FINE(0);
  /* ---- 3b. residual binary rows: Schur bits and their X-part source lists */
  int *lowrows = sc_buf(sc, SC_LOWROWS, sizeof(int) * (size_t)(nb ? nb : 1), 0);
  {
    int m = 0;
    for (int r = S + H; r < R; r++)
      if (row_pos[r] < 0) lowrows[m++] = r;
    for (int r = 0; r < S; r++)
      if (row_pos[r] < 0) lowrows[m++] = r;
  }
  const int nbw = (nb + 63) / 64;
  uint64_t *Sb = sc_buf(sc, SC_SB, (size_t)(nb ? nb : 1) * uw * sizeof(uint64_t), 1);
  uint64_t *Tb = sc_buf(sc, SC_TB, (size_t)(nb ? nb : 1) * (nbw ? nbw : 1) * sizeof(uint64_t), 1);
  int *xptr = sc_buf(sc, SC_XPTR, sizeof(int) * ((size_t)nb + 1), 0);
  int *xidx = sc_buf(sc, SC_XIDX, sizeof(int) * (size_t)(nnz ? nnz : 1), 0);
  OOM_CHECK();
  {
    int nx = 0;
    for (int m = 0; m < nb; m++) {
      int r = lowrows[m];
      uint64_t *s = Sb + (size_t)m * uw;
      xptr[m] = nx;
      for (int k = rptr[r]; k < rptr[r + 1]; k++) {
        const int q = cinfo[cidx[k]];
        if (q < 0) {
          bit_flip(s, ~q);
        } else {
          xidx[nx++] = q;
          bits_xor(s, G + (size_t)q * uw, uw);
        }
      }
      bit_flip(Tb + (size_t)m * nbw, m);
    }
    xptr[nb] = nx;
  }
  if (smem) { /* the residual rows accumulate in place as their terms become final, like the peeled ones */
    int *it = sc_buf(sc, SC_TMP, sizeof(int) * 2 * ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64), 0);
    int *rd = it + ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64);
    uint64_t *sortbuf = sc_buf(sc, SC_SORT, sizeof(uint64_t) * 2 * ((size_t)RQB_MAX_LT_DEGREE + (size_t)L + 64), 0);
    OOM_CHECK();
    for (int m = 0; m < nb; m++) {
      const int r = lowrows[m], cnt = xptr[m + 1] - xptr[m];
      for (int k = 0; k < cnt; k++) {
        it[k] = xidx[xptr[m] + k];
        rd[k] = level[it[k]];
      }
      ch.first[I + m] = ch.ng;
      chain_row(&ch, rd, it, cnt, r >= S + H && req->in_row[r - S - H] != RQB_ROW_NONE, maxlevel, sortbuf);
    }
    ch.first[I + nb] = ch.ng;
  }

  FINE(1);
  /* ---- 3c. HDPC Schur rows (H x U bytes) by the alpha recurrence.
   * HDPC[:,j] = alpha*HDPC[:,j+1] ^ e_b1(j) ^ e_b2(j), last column alpha^h
   * (lib/precode.c:60-83)  =>  sum_j HDPC[h][j] v_j = alpha^h y_{n-1} ^ sum_{j<=n-2, h in b(j)} y_j
   * with y_j = alpha*y_{j-1} ^ v_j.  Here v_j is the u-vector of column j: row q of G if the
   * column is peeled at q (its symbol is Y_q ^ G_q z), the unit vector if inactive. */
  const uint8_t *hb1 = bm->hb1, *hb2 = bm->hb2; /* the two HDPC rows of every column: per K', cached */
  const int uq = uw * 8; /* u64 words per byte-row of padded width 64*uw */
  uint8_t *Sh = sc_buf(sc, SC_SH, (size_t)H * (size_t)uq * 8, 1);
  uint64_t *ybuf = sc_buf(sc, SC_YBUF, (size_t)uq * 8, 1);
  uint64_t *accbuf = sc_buf(sc, SC_ACCBUF, (size_t)H * (size_t)uq * 8, 1);
  OOM_CHECK();
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

  FINE(2);
  /* ---- 3d. Gauss-Jordan over GF(2) on the binary Schur rows, transformation tracked */
  int *pivrow = sc_buf(sc, SC_PIVROW, sizeof(int) * (size_t)(U ? U : 1), 0); /* column t -> local row or -1 */
  int *pivcol_of_row = sc_buf(sc, SC_PIVCOL, sizeof(int) * (size_t)(nb ? nb : 1), 0);
  int *freecols = sc_buf(sc, SC_FREECOLS, sizeof(int) * (size_t)(U ? U : 1), 0);
  int rho = 0, nfree = 0;
  uint8_t *TQ = sc_buf(sc, SC_TQ, (size_t)H * (size_t)H, 1);
  int qrow_of_f[RQB_MAX_H];
  OOM_CHECK();
  /* on the device when that pays (rqb_planner.h): same pivots, same results */
  int on_device = 0;
  if (rqb_plan_usolve_hook && rqb_plan_usolve_mode != 1 && nb > 0 && (rqb_plan_usolve_mode == 2 || U >= 256)) {
    rqb_usolve_io uio = {nb, U, uw, nbw, H, (size_t)uq * 8, Sb, Tb, Sh, pivrow, TQ, qrow_of_f, 0, 0};
    const int ur = rqb_plan_usolve_hook(&uio);
    if (ur == 1) return 1; /* rank(A) < L */
    if (ur == 0) {
      on_device = 1;
      nfree = uio.nfree;
      rho = uio.rho;
      for (int t = 0, f = 0; t < U; t++)
        if (pivrow[t] < 0) freecols[f++] = t;
    }
  }
  if (!on_device) {
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
      if (nfree > H) return 1;
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
  FINE(3);
  /* ---- 3e. HDPC rows: eliminate pivot columns (beta = Sh[h][t]), then solve the
   *          H x nfree system Q over GF(256) with a tracked transformation TQ (H x H) */
  uint8_t *Q = sc_buf(sc, SC_Q, (size_t)H * (size_t)(nfree ? nfree : 1), 1);
  OOM_CHECK();
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
  {
    uint8_t used[RQB_MAX_H] = {0};
    for (int f = 0; f < nfree; f++) {
      int pr = -1;
      for (int h = 0; h < H; h++)
        if (!used[h] && Q[h * nfree + f]) {
          pr = h;
          break;
        }
      if (pr < 0) return 1; /* rank(A) < L */
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
  } /* !on_device */
  double t3 = now_s();

  FINE(4);
  /* ---- 4. emit the program.
   * Working rows: matrix row r -> WS row r; then r'_t (RP+t), z_t (Z+t), the scan
   * rows y_j, one row per part, then whatever the reduction trees need.
   * loc[s] = where the CURRENT value of working row s lives (an input row until
   * the first task writes it; NONE = known zero: nothing to read). */
  int NC = n / 32; /* scan chunks: ~32 entries each, few enough that combining them stays cheap */
  if (NC < 1) NC = 1;
  if (NC > 512) NC = 512;
  const uint32_t RP = (uint32_t)R;
  const uint32_t Z = RP + (uint32_t)U;
  const uint32_t YS = Z + (uint32_t)U;        /* y_j of the chunk-local alpha-scans, one row per column j < n */
  const uint32_t PT = YS + (uint32_t)n;
  const uint32_t WS_FIXED = PT + (uint32_t)nparts;
  /* arena layout: [IN | SYM | C | WS] */
  uint32_t row0[4];
  row0[RQB_SP_IN] = 0;
  row0[RQB_SP_SYM] = req->in_rows;
  row0[RQB_SP_C] = row0[RQB_SP_SYM] + req->sym_rows;
  const uint32_t zero_row = row0[RQB_SP_C] + (uint32_t)L;
  row0[RQB_SP_WS] = zero_row + 1;
  if (req->sym_rows < (uint32_t)req->n_out) return -1;
  if (req->out_row)
    for (int k = 0; k < req->n_out; k++)
      if (req->out_row[k] >= req->sym_rows) return -1;
  builder bd;
  memset(&bd, 0, sizeof(bd));
  bd.sc = sc;
  uint32_t smem_slots = 0, smem_tab_bits = 0;
  if (smem) {
    plan_view v;
    memset(&v, 0, sizeof(v));
    v.req = req; v.P = P; v.oh = oh; v.R = R; v.n = n; v.nlt = nlt; v.I = I; v.U = U; v.uw = uw; v.nb = nb;
    v.nbw = nbw; v.uq = uq; v.nfree = nfree; v.maxlevel = maxlevel;
    v.rptr = rptr; v.cidx = cidx; v.col_pos = col_pos; v.col_t = col_t; v.prow = prow; v.pcol = pcol; v.ucol = ucol;
    v.lowrows = lowrows; v.pivrow = pivrow; v.freecols = freecols; v.qrow_of_f = qrow_of_f;
    v.col_state = col_state; v.hb1 = hb1; v.hb2 = hb2; v.Sh = Sh; v.TQ = TQ; v.G = G; v.Sb = Sb; v.Tb = Tb;
    v.ch = &ch; v.slot_budget = slot_budget;
    memcpy(v.row0, row0, sizeof(row0));
    rc = emit_smem(&v, &bd, sc, &smem_slots, &smem_tab_bits);
    if (rc == -8) { /* the estimate was too optimistic for this block: HBM flavour */
      rqb_plan_request r2 = *req;
      r2.smem_budget = 0;
      return rqb_plan_build(&r2, out);
    }
    if (rc) return rc;
    bd.ws_next = 0; /* no working rows in HBM */
    goto emitted;
  }
  if ((uint64_t)row0[RQB_SP_WS] + (uint64_t)WS_FIXED + (uint64_t)nnz > RQB_MAX_ROWS) return -4;
  bd.ws_base = row0[RQB_SP_WS];
  bd.ws_next = WS_FIXED;
  uint32_t *loc = sc_buf(sc, SC_CURLOC, sizeof(uint32_t) * (size_t)WS_FIXED, 0);
  OOM_CHECK();
  for (uint32_t s = 0; s < WS_FIXED; s++) loc[s] = NONE_REF;
  for (int k = 0; k < nlt; k++)
    if (req->in_row[k] != RQB_ROW_NONE) {
      if (req->in_row[k] >= req->in_rows) return -1;
      loc[S + H + k] = row0[RQB_SP_IN] + req->in_row[k];
    }
  uint32_t *tmp = sc_buf(sc, SC_TMP, sizeof(uint32_t) * ((size_t)L + (size_t)nb + (size_t)4 * (size_t)n + (size_t)NC + 4096), 0);
  OOM_CHECK();
#define WSREF(s) (row0[RQB_SP_WS] + (uint32_t)(s))
#define PUSH(ns, ref)                                  \
  do {                                                 \
    uint32_t _r = (ref);                               \
    if (_r != NONE_REF) tmp[(ns)++] = RQB_SRC(_r, 1);  \
  } while (0)

  /* the triangular solve (phases A and E): levels base+1 .. base+maxlevel */
#define TRIANGULAR(base)                                                                        \
  do {                                                                                          \
    for (int p = 0; p < I; p++) {                                                               \
      for (int k = pfirst[p]; k < pfirst[p + 1]; k++) {                                         \
        uint32_t ns = 0;                                                                        \
        for (int e = pptr[k]; e < pptr[k + 1]; e++)                                             \
          PUSH(ns, pitems[e] >= 0 ? loc[prow[pitems[e]]] : loc[PT + (uint32_t)(-1 - pitems[e])]); \
        b_task(&bd, RQB_T_XOR, WSREF(PT + (uint32_t)k), 0, (base) + (uint32_t)plevel[k], tmp, ns); \
        loc[PT + (uint32_t)k] = WSREF(PT + (uint32_t)k);                                        \
      }                                                                                         \
      if (level[p] == 0) continue;                                                              \
      uint32_t ns = 0;                                                                          \
      PUSH(ns, loc[prow[p]]);                                                                   \
      for (int e = fptr[p]; e < fptr[p + 1]; e++)                                               \
        PUSH(ns, fitems[e] >= 0 ? loc[prow[fitems[e]]] : loc[PT + (uint32_t)(-1 - fitems[e])]); \
      b_task(&bd, RQB_T_XOR, WSREF(prow[p]), 0, (base) + (uint32_t)level[p], tmp, ns);          \
      loc[prow[p]] = WSREF(prow[p]);                                                            \
    }                                                                                           \
  } while (0)

  /* A: forward substitution Y = X^-1 b_top */
  TRIANGULAR(0u);
  uint32_t lv = (uint32_t)maxlevel + 1, end = lv;

  FINE(6);
  /* B: residual rows r_m = b_m ^ X_low*Y ; HDPC rows through NC chunk scans over columns 0..n-1 */
  for (int m = 0; m < nb; m++) {
    uint32_t ns = 0;
    if (xptr[m + 1] == xptr[m]) continue;
    PUSH(ns, loc[lowrows[m]]);
    for (int e = xptr[m]; e < xptr[m + 1]; e++) PUSH(ns, loc[prow[xidx[e]]]);
    uint32_t e2 = b_tree(&bd, RQB_T_XOR, WSREF(lowrows[m]), tmp, ns, lv);
    loc[lowrows[m]] = WSREF(lowrows[m]);
    if (e2 > end) end = e2;
  }
  int *cs = sc_buf(sc, SC_CS, sizeof(int) * ((size_t)NC + 1), 0);
  OOM_CHECK();
  for (int c = 0; c <= NC; c++) cs[c] = (int)((long)n * c / NC);
  for (int c = 0; c < NC; c++) {
    uint32_t ns = 0;
    for (int j = cs[c]; j < cs[c + 1]; j++) {
      tmp[ns++] = col_state[j] == 1 ? loc[prow[col_pos[j]]] : NONE_REF;
      loc[YS + (uint32_t)j] = WSREF(YS + (uint32_t)j);
    }
    b_task(&bd, RQB_T_SCAN, WSREF(YS + (uint32_t)cs[c]), 0, lv, tmp, ns);
  }
  lv = end + 1;
  end = lv;

  FINE(7);
  /* C1: r'_t = XOR_{m in Tb[pivrow t]} r_m ; HDPC base r_h = XOR_{j <= n-2, h in b(j)} y_j (chunk-local y) */
  for (int t = 0; t < U; t++) {
    if (pivrow[t] < 0) continue;
    const uint64_t *tb = Tb + (size_t)pivrow[t] * nbw;
    uint32_t ns = 0;
    for (int m = 0; m < nb; m++)
      if (bit_get(tb, m)) PUSH(ns, loc[lowrows[m]]);
    uint32_t e2 = b_tree(&bd, RQB_T_XOR, WSREF(RP + (uint32_t)t), tmp, ns, lv);
    loc[RP + (uint32_t)t] = WSREF(RP + (uint32_t)t);
    if (e2 > end) end = e2;
  }
  {
    /* bucket the columns by HDPC row in one pass: hcnt[h] = start of row h's list in tmp2 */
    uint32_t hcnt[RQB_MAX_H + 1];
    memset(hcnt, 0, sizeof(hcnt));
    for (int j = 0; j + 1 < n; j++) {
      hcnt[hb1[j] + 1]++;
      hcnt[hb2[j] + 1]++;
    }
    for (int h = 0; h < H; h++) hcnt[h + 1] += hcnt[h];
    uint32_t *tmp2 = tmp + (size_t)2 * (size_t)n + 64; /* tmp holds more than 4n words */
    uint32_t hcur[RQB_MAX_H];
    memcpy(hcur, hcnt, sizeof(hcur));
    for (int j = 0; j + 1 < n; j++) {
      uint32_t src = RQB_SRC(loc[YS + (uint32_t)j], 1);
      tmp2[hcur[hb1[j]]++] = src;
      tmp2[hcur[hb2[j]]++] = src;
    }
    for (int h = 0; h < H; h++) {
      uint32_t e2 = b_tree(&bd, RQB_T_XOR, WSREF(S + h), tmp2 + hcnt[h], hcnt[h + 1] - hcnt[h], lv);
      loc[S + h] = WSREF(S + h);
      if (e2 > end) end = e2;
    }
  }
  lv = end + 1;
  end = lv;

  FINE(8);
  /* C2: r'_h = r_h ^ sum_c Gc[h][c]*yend_c ^ sum_t beta[h][t]*r'_t  (GF leaves, XOR tree).
   * Gc[h][c'] = alpha^h alpha^(n-e_c') ^ sum_{c>c'} coef[c][h] alpha^(s_c-e_c'),
   * coef[c][h] = sum_{j in chunk c, j<=n-2, h in b(j)} alpha^(j-s_c+1). */
  {
    uint8_t *coef = sc_buf(sc, SC_COEF, (size_t)NC * (size_t)H, 1);
    for (int c = 0; c < NC; c++)
      for (int j = cs[c]; j < cs[c + 1] && j + 1 < n; j++) {
        uint8_t a = rqb_gf_pow2(&GF, j - cs[c] + 1);
        coef[c * H + hb1[j]] ^= a;
        coef[c * H + hb2[j]] ^= a;
      }
    /* suffix[h] = sum_{c > c1} coef[c][h] * alpha^(s_c - e_c1), built backwards in O(NC*H) */
    uint8_t suf[RQB_MAX_H];
    uint8_t *gc = sc_buf(sc, SC_CSLOT, (size_t)NC * (size_t)H + 64, 0);
    OOM_CHECK();
    memset(suf, 0, sizeof(suf));
    for (int c1 = NC - 1; c1 >= 0; c1--) {
      /* moving the reference point from e_{c1+1} (= s_{c1+2}) back to e_{c1} (= s_{c1+1}) multiplies by
       * alpha^(len of chunk c1+1) and brings in chunk c1+1 itself with exponent 0 */
      if (c1 + 1 < NC) {
        uint8_t a = rqb_gf_pow2(&GF, cs[c1 + 2] - cs[c1 + 1]);
        for (int h = 0; h < H; h++) suf[h] = rqb_gf_mul(&GF, suf[h], a) ^ coef[(c1 + 1) * H + h];
      }
      for (int h = 0; h < H; h++)
        gc[c1 * H + h] = rqb_gf_mul(&GF, rqb_gf_pow2(&GF, h), rqb_gf_pow2(&GF, n - cs[c1 + 1])) ^ suf[h];
    }
    for (int h = 0; h < H; h++) {
      const uint8_t *row = Sh + (size_t)h * (size_t)uq * 8;
      uint32_t ns = 0;
      for (int c1 = 0; c1 < NC; c1++) {
        uint8_t g = gc[c1 * H + h];
        uint32_t ref = loc[YS + (uint32_t)cs[c1 + 1] - 1]; /* y at the end of chunk c1 */
        if (g && ref != NONE_REF) tmp[ns++] = RQB_SRC(ref, g);
      }
      for (int t = 0; t < U; t++)
        if (pivrow[t] >= 0 && row[t] && loc[RP + (uint32_t)t] != NONE_REF) tmp[ns++] = RQB_SRC(loc[RP + (uint32_t)t], row[t]);
      if (loc[S + h] != NONE_REF) tmp[ns++] = RQB_SRC(loc[S + h], 1);
      uint32_t e2 = b_tree(&bd, RQB_T_GF, WSREF(S + h), tmp, ns, lv);
      loc[S + h] = WSREF(S + h);
      if (e2 > end) end = e2;
    }
  }
  lv = end + 1;
  end = lv;
  FINE(9);
  /* C3: z_f = sum_h TQ[qrow(f)][h] * r'_h */
  for (int f = 0; f < nfree; f++) {
    uint32_t ns = 0;
    for (int h = 0; h < H; h++) {
      uint8_t b = TQ[qrow_of_f[f] * H + h];
      if (b && loc[S + h] != NONE_REF) tmp[ns++] = RQB_SRC(loc[S + h], b);
    }
    uint32_t e2 = b_tree(&bd, RQB_T_GF, WSREF(Z + (uint32_t)freecols[f]), tmp, ns, lv);
    loc[Z + (uint32_t)freecols[f]] = WSREF(Z + (uint32_t)freecols[f]);
    if (e2 > end) end = e2;
  }
  lv = end + 1;
  end = lv;
  /* C4: z_t = r'_t ^ XOR_{f: bit} z_f for pivot columns */
  for (int t = 0; t < U; t++) {
    if (pivrow[t] < 0) continue;
    const uint64_t *ps = Sb + (size_t)pivrow[t] * uw;
    uint32_t ns = 0;
    PUSH(ns, loc[RP + (uint32_t)t]);
    for (int f = 0; f < nfree; f++)
      if (bit_get(ps, freecols[f])) PUSH(ns, loc[Z + (uint32_t)freecols[f]]);
    uint32_t e2 = b_tree(&bd, RQB_T_XOR, WSREF(Z + (uint32_t)t), tmp, ns, lv);
    loc[Z + (uint32_t)t] = WSREF(Z + (uint32_t)t);
    if (e2 > end) end = e2;
  }
  lv = end + 1;
  end = lv;

  FINE(10);
  /* Back-substitution into the peeled part, x = X^-1 (b_top ^ U_top z), one of two ways.
   *
   * F ("four Russians"): x_p = Y_p ^ (G z)_p with Y from phase A and G = X^-1 U_top, the
   * bit matrix the host already holds.  G is dense (about a third of its bits are set), so
   * the u inactive symbols are taken 8 at a time: table rows T_j[m] = XOR of the z of group j
   * selected by the byte m (built from two 4-bit half tables, 2 levels), and every x_p is ONE
   * table-gather task (RQB_T_TAB) of at most ceil(u/8) table rows.  Costs more sources than a second triangular
   * solve but needs 4 levels instead of ~500 -- the solve kernel is bound by the latency of
   * its dependency levels, not by bytes.
   *
   * D+E: patch b_top with U_top z and run the triangular solve again; fewer bytes, used when
   * u is so large (K' in the tens of thousands) that F's gathers would dominate. */
  const int ng = (U + 7) / 8;
  pthread_once(&backsub_once, backsub_init);
  const int fr_mode = g_fr_mode;
  const int use_fr = I > 0 && fr_mode != 2 && (fr_mode == 1 || (size_t)ng * (size_t)I * 2 <= (size_t)nnz * 5);
  if (use_fr) {
    /* table rows sit at fixed positions, tab_base + 256*j + m, so that a TAB task needs
     * nothing but the bytes of its row of G; only the entries some row uses (and the two
     * half-table entries they are made of) are computed */
    uint8_t *used8 = sc_buf(sc, SC_FRUSED, (size_t)ng * 256 + 64, 1);
    /* a decoder only back-substitutes the peeled rows whose symbol is a term of one of its outputs
     * (at 10 % loss about two thirds of them) */
    uint8_t *needed = sc_buf(sc, SC_NEEDED, (size_t)L + 8, 0);
    OOM_CHECK();
    if (req->want_c) {
      memset(needed, 1, (size_t)L);
    } else {
      memset(needed, 0, (size_t)L);
      for (int k = 0; k < req->n_out; k++) {
        uint32_t idx[RQB_MAX_LT_DEGREE];
        int cnt = rqb_host_lt_indices(&P, req->out_isi[k], idx);
        for (int q = 0; q < cnt; q++) needed[idx[q]] = 1;
      }
    }
    const uint32_t tab0 = bd.ws_next;
    bd.tab_base = bd.ws_base + tab0;
    bd.ws_next += (uint32_t)ng * 256u;
    for (int p = 0; p < I; p++) {
      if (!needed[pcol[p]]) continue;
      const uint8_t *g = (const uint8_t *)(G + (size_t)p * uw);
      for (int j = 0; j < ng; j++) used8[(size_t)j * 256 + g[j]] = 1;
    }
    for (int j = 0; j < ng; j++) {
      uint8_t *use = used8 + (size_t)j * 256;
      for (int m = 1; m < 256; m++)
        if (use[m] == 1 && (m & 15) && (m >> 4)) { /* made of the half-table entries m&15 and m&0xf0 */
          if (!use[m & 15]) use[m & 15] = 2;
          if (!use[m & 0xf0]) use[m & 0xf0] = 2;
        }
      for (int m = 1; m < 256; m++) {
        if (!use[m]) continue;
        const uint32_t row = bd.tab_base + (uint32_t)j * 256u + (uint32_t)m;
        if ((m & 15) && (m >> 4)) {
          uint32_t pair[2] = {bd.tab_base + (uint32_t)j * 256u + (uint32_t)(m & 15),
                              bd.tab_base + (uint32_t)j * 256u + (uint32_t)(m & 0xf0)};
          b_task(&bd, RQB_T_XOR, row, 0, lv + 1, pair, 2);
        } else { /* a half-table entry: up to 4 of the z themselves */
          uint32_t ns = 0;
          for (int bit = 0; bit < 8; bit++)
            if (m >> bit & 1) {
              int t = 8 * j + bit;
              if (t < U) PUSH(ns, loc[Z + (uint32_t)t]);
            }
          b_task(&bd, RQB_T_XOR, row, 0, lv, tmp, ns);
        }
      }
    }
    end = lv + 2;
    for (int p = 0; p < I; p++) {
      if (!needed[pcol[p]]) continue; /* nobody reads this x_p */
      const uint8_t *g = (const uint8_t *)(G + (size_t)p * uw);
      int any = 0;
      for (int j = 0; j < ng; j++) any |= g[j];
      if (!any) continue; /* x_p = Y_p */
      /* an encoder keeps C: x_p is the intermediate symbol of column pcol[p], written straight to
       * its row of the C space (no copy task later) */
      const uint32_t xdst = req->want_c ? row0[RQB_SP_C] + (uint32_t)pcol[p] : WSREF(prow[p]);
      b_tab(&bd, xdst, loc[prow[p]] != NONE_REF ? loc[prow[p]] : zero_row, lv + 2, g, (uint32_t)ng);
      loc[prow[p]] = xdst;
    }
    FINE(11);
    lv = end + 1;
    end = lv;
  } else {
    /* D: b_top' = b_top ^ U_top z.  A row without inactive columns just goes back to
     * its input row (no task); the others are rebuilt from the input row. */
    for (int p = 0; p < I; p++) {
      int r = prow[p];
      uint32_t ns = 0;
      uint32_t orig = r >= S + H ? (req->in_row[r - S - H] != RQB_ROW_NONE ? row0[RQB_SP_IN] + req->in_row[r - S - H] : NONE_REF)
                                 : NONE_REF;
      for (int k = rptr[r]; k < rptr[r + 1]; k++)
        if (col_state[cidx[k]] == 2) PUSH(ns, loc[Z + (uint32_t)col_t[cidx[k]]]);
      if (ns == 0) {
        loc[r] = orig;
        continue;
      }
      PUSH(ns, orig);
      uint32_t e2 = b_tree(&bd, RQB_T_XOR, WSREF(r), tmp, ns, lv);
      loc[r] = WSREF(r);
      if (e2 > end) end = e2;
    }
    FINE(11);
    /* E: x = X^-1 b_top' */
    TRIANGULAR(end);
    lv = end + (uint32_t)maxlevel + 1;
    end = lv;
  }
  FINE(12);
  /* O: outputs.  C[col] sits in the row that pivoted on col, or in z. */
  uint32_t *cloc = sc_buf(sc, SC_CSLOT, sizeof(uint32_t) * (size_t)L, 0);
  OOM_CHECK();
  for (int c = 0; c < L; c++) cloc[c] = col_state[c] == 1 ? loc[prow[col_pos[c]]] : loc[Z + (uint32_t)col_t[c]];
  if (req->want_c)
    for (int c = 0; c < L; c++) {
      uint32_t ns = 0;
      if (cloc[c] == row0[RQB_SP_C] + (uint32_t)c) continue; /* already written in place */
      PUSH(ns, cloc[c]);
      b_task(&bd, RQB_T_XOR, row0[RQB_SP_C] + (uint32_t)c, 0, lv, tmp, ns);
    }
  for (int k = 0; k < req->n_out; k++) {
    uint32_t idx[RQB_MAX_LT_DEGREE];
    int cnt = rqb_host_lt_indices(&P, req->out_isi[k], idx);
    uint32_t ns = 0;
    for (int q = 0; q < cnt; q++) PUSH(ns, cloc[idx[q]]);
    b_tree(&bd, RQB_T_XOR, row0[RQB_SP_SYM] + (req->out_row ? req->out_row[k] : (uint32_t)k), tmp, ns, lv);
  }
#undef PUSH
#undef WSREF
#undef TRIANGULAR
#undef OOM_CHECK_UNUSED
  if ((uint64_t)row0[RQB_SP_WS] + bd.ws_next > RQB_MAX_ROWS) return -4;

emitted:
  FINE(13);
  OOM_CHECK();
  rqb_plan *plan = plan_acquire();
  if (!plan) return -7;
  size_t tot_levels = 0;
  /* the shared-memory flavour pads its XOR lists with slot 0 */
  rc = write_pages(&bd, plan, smem ? 0u : zero_row, smem, &tot_levels, req->pages_buf, req->pages_buf_cap);
  if (rc) {
    rqb_plan_free(plan);
    return rc;
  }
  double t4 = now_s();
  FINE(14);

  plan->P = P;
  plan->K = req->K;
  plan->overhead = oh;
  plan->n_ws_rows = bd.ws_next;
  memcpy(plan->row0, row0, sizeof(row0));
  plan->n_rows = row0[RQB_SP_WS] + bd.ws_next;
  plan->zero_row = zero_row;
  plan->n_c_rows = req->want_c ? (uint32_t)L : 0;
  plan->n_out = (uint32_t)req->n_out;
  plan->smem = smem;
  plan->slice_bytes = slice;
  plan->n_slots = smem_slots;
  plan->tab_bits = smem_tab_bits;
  plan->st.i = I;
  plan->st.u = U;
  plan->st.nb = nb;
  plan->st.rho = rho;
  plan->st.nfree = nfree;
  plan->st.levels_fwd = maxlevel + 1;
  plan->st.n_parts = nparts;
  plan->st.n_levels = (int)tot_levels;
  plan->st.n_tasks = (int)bd.nt;
  plan->st.n_pages = (int)plan->n_pages;
  plan->st.n_srcs = bd.tot_x;
  plan->st.n_gf_srcs = bd.tot_gf;
  plan->st.n_horner = bd.tot_h;
  plan->st.nnz = (size_t)nnz;
  plan->st.t_matrix = t1 - t0;
  plan->st.t_peel = t2 - t1;
  plan->st.t_dense = t3 - t2;
  plan->st.t_emit = t4 - t3;
  *out = plan;
  return 0;
}


/* ------------------------------------------- program from a reference schedule
 * Turns an already ordered sequence of reference-format row operations
 * (sched_op, include/sched.h:6-10, in the order precode_matrix_apply_sched applies
 * them, lib/precode.c:23-32) followed by a row gather (the two permutations of
 * precode_matrix_intermediate, lib/precode.c:379-389) into ONE device program for
 * the solve kernel, instead of one launch per dependency level:
 *   - ops are levelised by their true dependencies (an op needs the last write of
 *     its source and of its destination; it must not overtake a pending read of
 *     its destination);
 *   - successive accumulations into one destination whose sources were all ready
 *     in time are MERGED into one gather task of up to 8 sources (the reference
 *     reads and writes the destination once per op) -- legal because GF(256)
 *     addition commutes and nobody reads the destination in between;
 *   - the gather runs as a last level into a second half of the arena.
 * Rows: the matrix occupies arena rows [base, base+nrows), the gathered result goes
 * to [out_base, out_base+nrows); zero_row is an all-zero row (XOR list padding). */
typedef struct {
  uint32_t dst, level;
  uint8_t n, gf, closed;
  uint32_t src[RQB_MAX_SRCS]; /* row | beta << 24 */
} stask;

int rqb_plan_from_schedule(const void *ops_v, size_t nops, uint32_t nrows, const uint32_t *gather_map, uint32_t base,
                           uint32_t out_base, uint32_t zero_row, rqb_plan **out) {
  pthread_once(&tables_once, tables_build);
  const rqb_rowop *ops = (const rqb_rowop *)ops_v;
  *out = NULL;
  if ((uint64_t)out_base + nrows > RQB_MAX_ROWS || (uint64_t)base + nrows > RQB_MAX_ROWS) return -4;
  stask *t = malloc(sizeof(stask) * (nops ? nops : 1));
  uint32_t *lw = calloc((size_t)nrows + 1, 4), *lr = calloc((size_t)nrows + 1, 4);
  int64_t *open = malloc(sizeof(int64_t) * ((size_t)nrows + 1));
  if (!t || !lw || !lr || !open) {
    free(t);
    free(lw);
    free(lr);
    free(open);
    return -7;
  }
  for (uint32_t r = 0; r < nrows; r++) open[r] = -1;
  size_t nt = 0;
  uint32_t maxlevel = 0;
  int rc = 0;
  for (size_t q = 0; q < nops && !rc; q++) {
    const uint32_t i = ops[q].i, j = ops[q].j;
    uint32_t beta = ops[q].beta;
    if (i >= nrows || (beta && j >= nrows)) {
      rc = -1;
      break;
    }
    if (beta == 0 || i == j) { /* oscal (multiplier in j; < 2 is a no-op, oblas_avx.c:94-95) or a row added to itself */
      uint32_t mult = beta == 0 ? (j & 0xffu) : (beta ^ 1u); /* x ^ b*x = (1^b)*x */
      if (beta == 0 && mult < 2) continue;
      if (open[i] >= 0) t[open[i]].closed = 1;
      open[i] = -1;
      uint32_t lvl = (lw[i] > lr[i] ? lw[i] : lr[i]) + 1;
      stask *k = &t[nt++];
      k->dst = i; k->level = lvl; k->n = 1; k->gf = 1; k->closed = 1;
      k->src[0] = RQB_SRC(base + i, mult);
      lw[i] = lvl;
      if (lvl > maxlevel) maxlevel = lvl;
      continue;
    }
    if (open[j] >= 0) t[open[j]].closed = 1; /* j is read: later accumulations into j start a new task */
    open[j] = -1;
    int64_t o = open[i];
    if (o >= 0 && !t[o].closed && lw[j] < t[o].level && t[o].n < RQB_MAX_SRCS) {
      stask *k = &t[o];
      k->src[k->n++] = RQB_SRC(base + j, beta);
      if (beta != 1) k->gf = 1;
      if (lr[j] < k->level) lr[j] = k->level;
      continue;
    }
    uint32_t lvl = lw[i] > lw[j] ? lw[i] : lw[j];
    if (lr[i] > lvl) lvl = lr[i];
    lvl++;
    if (o >= 0) t[o].closed = 1;
    stask *k = &t[nt];
    k->dst = i; k->level = lvl; k->n = 2; k->gf = beta != 1; k->closed = 0;
    k->src[0] = RQB_SRC(base + i, 1);
    k->src[1] = RQB_SRC(base + j, beta);
    open[i] = (int64_t)nt++;
    lw[i] = lvl;
    if (lr[j] < lvl) lr[j] = lvl;
    if (lvl > maxlevel) maxlevel = lvl;
  }
  if (!rc) {
    scratch_t *sc = sc_get();
    builder bd;
    memset(&bd, 0, sizeof(bd));
    bd.sc = sc;
    if (!sc) rc = -7;
    else sc->oom = 0;
    for (size_t k = 0; k < nt && !rc; k++) b_task(&bd, t[k].gf ? RQB_T_GF : RQB_T_XOR, base + t[k].dst, 0, t[k].level, t[k].src, t[k].n);
    for (uint32_t k = 0; k < nrows && !rc; k++) {
      uint32_t src = RQB_SRC(base + gather_map[k], 1);
      if (gather_map[k] >= nrows) {
        rc = -1;
        break;
      }
      b_task(&bd, RQB_T_XOR, out_base + k, 0, maxlevel + 1, &src, 1);
    }
    if (!rc && sc->oom) rc = -7;
    rqb_plan *plan = rc ? NULL : plan_acquire();
    if (!rc && !plan) rc = -7;
    if (!rc) {
      size_t tot_levels = 0;
      rc = write_pages(&bd, plan, zero_row, 0, &tot_levels, NULL, 0);
      if (rc) {
        rqb_plan_free(plan);
      } else {
        memset(&plan->st, 0, sizeof(plan->st));
        plan->st.n_levels = (int)tot_levels;
        plan->st.n_tasks = (int)bd.nt;
        plan->st.n_pages = (int)plan->n_pages;
        plan->st.n_srcs = bd.tot_x;
        plan->st.n_gf_srcs = bd.tot_gf;
        plan->n_ws_rows = 0;
        plan->smem = 0;
        plan->slice_bytes = plan->n_slots = plan->tab_bits = 0;
        plan->n_rows = out_base + nrows;
        plan->zero_row = zero_row;
        *out = plan;
      }
    }
  }
  free(t);
  free(lw);
  free(lr);
  free(open);
  return rc;
}
