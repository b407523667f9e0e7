/* rqb_solver.c -- host side of the rqb200.h C ABI: per-block device contexts,
 * buffer pool, process-wide cache of encoder plans, batched row ops and the
 * reference-format schedule replay.  Plain C; the device work is done by the
 * kernels behind rqb_device.h.  There is no CPU fallback anywhere in this
 * file: symbol bytes are only ever combined on the GPU. */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rqb200.h"
#include "rqb_device.h"
#include "rqb_planner.h"
#include "rqb_program.h"

static __thread char g_err[320];

const char *rqb_last_error(void) { return g_err; }

static int dev_fail(int e, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s failed: %s", what, rqb_dev_last_error());
  (void)e;
  return RQB_E_NODEVICE;
}
#define DEV(x)                               \
  do {                                       \
    int _e = (x);                            \
    if (_e) return dev_fail(_e, #x);         \
  } while (0)

/* ------------------------------------------------ host time accounting */
#include "rqb_hostcopy.h"
#include "rqb_prof.h"
static _Atomic unsigned long long g_prof_ns[RQB_PF_COUNT];
static int g_prof_on;
static pthread_once_t g_prof_once = PTHREAD_ONCE_INIT;
static const char *const g_prof_names[RQB_PF_COUNT] = {
    "gen.load", "gen.upload", "gen.plan", "gen.run", "gen.sync", "emit.source", "emit.window", "add.create",
    "add.copy", "add.write", "repair.upload", "repair.request", "repair.plan", "repair.pages", "repair.args",
    "repair.run", "repair.fetch", "repair.write", "free", "range.generate", "range.queue", "range.wait",
    "add_symbols"};
static void prof_init(void) {
  const char *e = getenv("NANORQ_B200_PROFILE");
  g_prof_on = e && e[0] == '1';
}
int rqb_prof_enabled(void) {
  pthread_once(&g_prof_once, prof_init);
  return g_prof_on;
}
double rqb_prof_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void rqb_prof_add(int slot, double seconds) {
  if (slot >= 0 && slot < RQB_PF_COUNT) g_prof_ns[slot] += (unsigned long long)(seconds * 1e9);
}
void rqb_host_profile(double *seconds, int n) {
  for (int k = 0; k < n; k++) seconds[k] = k < RQB_PF_COUNT ? 1e-9 * (double)g_prof_ns[k] : 0.0;
}
const char *rqb_host_profile_name(int k) { return k >= 0 && k < RQB_PF_COUNT ? g_prof_names[k] : NULL; }
void rqb_host_profile_reset(void) {
  for (int k = 0; k < RQB_PF_COUNT; k++) g_prof_ns[k] = 0;
}

int rqb_device_count(void) { return rqb_dev_count(); }
int rqb_set_device(int dev) {
  DEV(rqb_dev_set(dev));
  rqb_dev_set_default(dev);
  return 0;
}
unsigned long long rqb_kernel_launches(void) { return rqb_dev_launch_count(); }
void rqb_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h) { rqb_dev_transfer_bytes(h2d, d2h); }

/* CUDA's current device is per thread; contexts remember theirs and every entry
 * point binds the calling thread to it (worker threads start on device 0) */
static int bind_dev(int dev) {
  if (rqb_dev_get() == dev) return 0;
  return rqb_dev_set(dev);
}
#define BIND(dev) DEV(bind_dev(dev))

int rqb_block_params_init(int K, rqb_block_params *out) {
  rqb_params P;
  if (rqb_params_init(K, &P)) return RQB_E_ARG;
  memcpy(out, &P, sizeof(P));
  return 0;
}

int rqb_lt_row_indices(int K, uint32_t isi, uint32_t *out) {
  rqb_params P;
  if (rqb_params_init(K, &P)) return RQB_E_ARG;
  return rqb_host_lt_indices(&P, isi, out);
}

#define RQB_ARGS_BYTES 16384 /* room for 256 blocks in one batched launch */

/* NANORQ_B200_FLAVOUR = hbm forces the HBM flavour of the solve program (experiments, tests);
 * default: the shared-memory flavour whenever a block's rows fit (rqb_program.h) */
static uint32_t g_smem_budget = RQB_SMEM_BUDGET_BYTES;
static pthread_once_t g_flavour_once = PTHREAD_ONCE_INIT;
static void flavour_init(void) {
  const char *e = getenv("NANORQ_B200_FLAVOUR");
  if (e && !strcmp(e, "hbm")) g_smem_budget = 0;
}
static uint32_t solver_smem_budget(void) {
  pthread_once(&g_flavour_once, flavour_init);
  return g_smem_budget;
}
static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

/* ------------------------------------------------------------ buffer pool
 * cudaMalloc / cudaMallocHost cost milliseconds; blocks come and go at wire
 * rate, so freed buffers are kept and handed out again (per device). */
typedef struct pool_ent {
  void *p;
  size_t bytes;
  int dev, pinned;
  struct pool_ent *next;
} pool_ent;
static pool_ent *g_pool;
static pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;
/* bytes parked in the pool and in recycled solver contexts, and their limit
 * (rqb_set_cache_limit / NANORQ_B200_CACHE_MB; default 48 GiB of the B200's 180: sixteen threads
 * working on maximum-size blocks, K = 56403, keep 10 GiB of contexts in rotation, and a limit below
 * the working set turns every block into cudaMalloc + cudaFree).  What would go over the limit is
 * handed back to the driver instead of being kept. */
static _Atomic size_t g_pool_bytes, g_shell_bytes;
static _Atomic size_t g_cache_limit = (size_t)48 << 30;
static pthread_once_t g_cache_once = PTHREAD_ONCE_INIT;
static void cache_limit_init(void) {
  const char *e = getenv("NANORQ_B200_CACHE_MB");
  if (e && *e) g_cache_limit = (size_t)strtoull(e, NULL, 10) << 20;
}
static size_t cache_limit(void) {
  pthread_once(&g_cache_once, cache_limit_init);
  return g_cache_limit;
}
void rqb_set_cache_limit(size_t bytes) {
  pthread_once(&g_cache_once, cache_limit_init);
  g_cache_limit = bytes;
}
void rqb_cache_stats(size_t *cached_bytes, size_t *limit_bytes) {
  if (cached_bytes) *cached_bytes = g_pool_bytes + g_shell_bytes;
  if (limit_bytes) *limit_bytes = cache_limit();
}
int rqb_device_mem_info(size_t *free_bytes, size_t *total_bytes) {
  size_t f = 0, t = 0;
  if (rqb_dev_count() <= 0) return RQB_E_NODEVICE;
  if (rqb_dev_get() != rqb_dev_default() && rqb_dev_set(rqb_dev_default())) return RQB_E_NODEVICE;
  if (rqb_dev_mem_info(&f, &t)) return RQB_E_NODEVICE;
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}

static size_t pool_class(size_t bytes) {
  size_t c = 4096;
  while (c < bytes && c < ((size_t)1 << 20)) c <<= 1;
  if (c >= bytes) return c;
  return round_up(bytes, (size_t)1 << 20);
}

/* slow-path events (each one takes the driver's global lock for milliseconds): fresh
 * pinned / device allocations, arena regrowths, contexts created.  In steady state all
 * of them must stay flat; rqb_slow_path_counters lets a benchmark check that. */
static _Atomic unsigned long long g_cnt_alloc_pinned, g_cnt_alloc_dev, g_cnt_regrow, g_cnt_ctx_new;
void rqb_slow_path_counters(unsigned long long out[4]) {
  out[0] = g_cnt_alloc_pinned;
  out[1] = g_cnt_alloc_dev;
  out[2] = g_cnt_regrow;
  out[3] = g_cnt_ctx_new;
}

static int pool_get(void **out, size_t bytes, int pinned) {
  size_t cls = pool_class(bytes);
  int dev = pinned ? -1 : rqb_dev_get();
  pthread_mutex_lock(&g_pool_mu);
  for (pool_ent **pp = &g_pool; *pp; pp = &(*pp)->next) {
    pool_ent *e = *pp;
    if (e->pinned == pinned && e->dev == dev && e->bytes == cls) {
      *pp = e->next;
      pthread_mutex_unlock(&g_pool_mu);
      g_pool_bytes -= e->bytes;
      *out = e->p;
      free(e);
      return 0;
    }
  }
  pthread_mutex_unlock(&g_pool_mu);
  if (pinned) g_cnt_alloc_pinned++; else g_cnt_alloc_dev++;
  return pinned ? rqb_host_malloc(out, cls) : rqb_dev_malloc(out, cls);
}

/* the calling thread must be bound to the buffer's device */
static void buf_free(void *p, int pinned) {
  if (!p) return;
  if (pinned) rqb_host_free(p); else rqb_dev_free(p);
}

static void pool_put(void *p, size_t bytes, int pinned) {
  if (!p) return;
  const size_t cls = pool_class(bytes);
  pool_ent *e = g_pool_bytes + g_shell_bytes + cls > cache_limit() ? NULL : malloc(sizeof(*e));
  if (!e) { /* over the limit (or out of memory): really free */
    buf_free(p, pinned);
    return;
  }
  e->p = p;
  e->bytes = cls;
  e->pinned = pinned;
  e->dev = pinned ? -1 : rqb_dev_get();
  g_pool_bytes += cls;
  pthread_mutex_lock(&g_pool_mu);
  e->next = g_pool;
  g_pool = e;
  pthread_mutex_unlock(&g_pool_mu);
}

static void pool_drain(void) { /* hand every pooled buffer back to the driver */
  pthread_mutex_lock(&g_pool_mu);
  pool_ent *e = g_pool;
  g_pool = NULL;
  pthread_mutex_unlock(&g_pool_mu);
  const int cur = rqb_dev_get();
  while (e) {
    pool_ent *n = e->next;
    if (!e->pinned && e->dev >= 0 && rqb_dev_get() != e->dev) rqb_dev_set(e->dev);
    buf_free(e->p, e->pinned);
    g_pool_bytes -= e->bytes;
    free(e);
    e = n;
  }
  if (cur >= 0 && rqb_dev_get() != cur) rqb_dev_set(cur);
}

/* --------------------------------------------------- encoder plan cache
 * The constraint matrix of an encoder depends only on K (isi = identity), like
 * the schedule nanorq_precalculate keeps in rq->S (lib/nanorq.c:393-401); here
 * it is cached process-wide together with its device copy. */
typedef struct enc_plan {
  int K, Kparams, want_c, dev, hbm; /* hbm: asked for the HBM flavour */
  uint32_t n_out, in_rows, sym_rows; /* the arena layout is part of the program */
  rqb_plan *plan;
  uint8_t *d_pages;
  int refs;               /* solver contexts whose current program this is (parked ones included) */
  unsigned long long use; /* LRU stamp */
  struct enc_plan *next;
} enc_plan;
static enc_plan *g_enc_plans;
static pthread_mutex_t g_plan_mu = PTHREAD_MUTEX_INITIALIZER;
static unsigned long long g_enc_clock;
#define RQB_ENC_PLANS_MAX 48 /* cached encoder programs (one per K, device and window size in use) */

static void enc_plan_free(enc_plan *e) { /* caller holds g_plan_mu; e is unlinked and unreferenced */
  const int cur = rqb_dev_get();
  if (cur != e->dev) rqb_dev_set(e->dev);
  if (e->d_pages) rqb_dev_free(e->d_pages);
  if (cur >= 0 && cur != e->dev) rqb_dev_set(cur);
  rqb_plan_free(e->plan);
  free(e);
}
static void enc_plan_unref(enc_plan *e) {
  if (!e) return;
  pthread_mutex_lock(&g_plan_mu);
  e->refs--;
  pthread_mutex_unlock(&g_plan_mu);
}
/* caller holds g_plan_mu: drop least recently used entries nobody runs until the cache is within `keep` */
static void enc_plans_trim(int keep) {
  for (;;) {
    int n = 0;
    enc_plan **victim = NULL;
    for (enc_plan **pp = &g_enc_plans; *pp; pp = &(*pp)->next) {
      n++;
      if ((*pp)->refs == 0 && (!victim || (*pp)->use < (*victim)->use)) victim = pp;
    }
    if (n <= keep || !victim) return;
    enc_plan *e = *victim;
    *victim = e->next;
    enc_plan_free(e);
  }
}

/* ------------------------------------------------- U-solve on the device
 * rqb_plan_build's hook (rqb_planner.h): the Schur elimination of one block as ONE kernel launch
 * on the calling thread's current device.  Every planning thread keeps a small context of its own
 * (stream, pinned and device buffer, grown on demand) for the lifetime of the thread. */
typedef struct {
  void *stream;
  uint8_t *h_buf, *d_buf;
  size_t cap;
  int dev;
} usolve_ctx;
static pthread_key_t g_usolve_key;
static pthread_once_t g_usolve_once = PTHREAD_ONCE_INIT;
static void usolve_ctx_free(void *v) {
  usolve_ctx *c = v;
  if (!c) return;
  if (rqb_dev_get() != c->dev) rqb_dev_set(c->dev);
  if (c->stream) rqb_stream_sync(c->stream);
  if (c->h_buf) rqb_host_free(c->h_buf);
  if (c->d_buf) rqb_dev_free(c->d_buf);
  if (c->stream) rqb_stream_destroy(c->stream);
  free(c);
}
static void usolve_make_key(void) { pthread_key_create(&g_usolve_key, usolve_ctx_free); }

static int usolve_on_device(rqb_usolve_io *io) {
  const int dev = rqb_dev_get();
  if (dev < 0) return -1;
  pthread_once(&g_usolve_once, usolve_make_key);
  usolve_ctx *c = pthread_getspecific(g_usolve_key);
  if (c && c->dev != dev) { /* the thread moved to another device: start over there */
    usolve_ctx_free(c);
    c = NULL;
    pthread_setspecific(g_usolve_key, NULL);
  }
  if (!c) {
    c = calloc(1, sizeof(*c));
    if (!c) return -1;
    c->dev = dev;
    if (rqb_stream_create(&c->stream)) {
      free(c);
      return -1;
    }
    pthread_setspecific(g_usolve_key, c);
  }
  const size_t hdr = rqb_usolve_header_bytes();
  const size_t need = rqb_usolve_buffer_bytes(io->nb, io->U, io->uw, io->nbw, io->H, (int)io->sh_stride);
  if (need > c->cap) {
    if (c->h_buf) rqb_host_free(c->h_buf);
    if (c->d_buf) rqb_dev_free(c->d_buf);
    c->h_buf = c->d_buf = NULL;
    c->cap = 0;
    const size_t cap = need + need / 2;
    g_cnt_alloc_pinned++;
    g_cnt_alloc_dev++;
    if (rqb_host_malloc((void **)&c->h_buf, cap) || rqb_dev_malloc((void **)&c->d_buf, cap)) return -1;
    c->cap = cap;
  }
  /* pack: header | Sb | Tb (device fills it) | Sh | pivrow (device fills it) */
  int32_t *h = (int32_t *)c->h_buf;
  memset(c->h_buf, 0, hdr);
  h[0] = io->nb; h[1] = io->U; h[2] = io->uw; h[3] = io->nbw; h[4] = io->H; h[5] = (int32_t)io->sh_stride;
  uint8_t *p_sb = c->h_buf + hdr;
  uint8_t *p_tb = p_sb + (size_t)io->nb * io->uw * 8;
  uint8_t *p_sh = p_tb + (size_t)io->nb * io->nbw * 8;
  uint8_t *p_piv = p_sh + (((size_t)io->H * io->sh_stride + 15) & ~(size_t)15);
  memcpy(p_sb, io->Sb, (size_t)io->nb * io->uw * 8);
  memcpy(p_sh, io->Sh, (size_t)io->H * io->sh_stride);
  const size_t in_bytes = (size_t)(p_piv - c->h_buf);
  int e = rqb_copy_h2d(c->d_buf, c->h_buf, in_bytes, c->stream);
  e = e ? e : rqb_launch_usolve(c->d_buf, io->nb, io->U, io->uw, io->nbw, c->stream);
  if (e) return -1; /* e.g. the rows do not fit a CTA's shared memory: the host code runs */
  e = rqb_copy_d2h(c->h_buf, c->d_buf, need, c->stream);
  e = e ? e : rqb_stream_sync(c->stream);
  if (e) return -1;
  if (h[6] != 0) return 1; /* singular */
  io->nfree = h[7];
  io->rho = h[8];
  memcpy(io->qrow_of_f, h + 10, sizeof(int) * 16 < sizeof(int) * (size_t)io->H ? sizeof(int) * 16 : sizeof(int) * (size_t)io->H);
  memcpy(io->TQ, c->h_buf + 10 * 4 + 16 * 4, (size_t)io->H * io->H);
  memcpy(io->Sb, p_sb, (size_t)io->nb * io->uw * 8);
  memcpy(io->Tb, p_tb, (size_t)io->nb * io->nbw * 8);
  memcpy(io->pivrow, p_piv, (size_t)io->U * 4);
  return 0;
}

void rqb_set_usolve_mode(int mode) { rqb_plan_usolve_mode = mode >= 0 && mode <= 2 ? mode : 0; }

/* ------------------------------------------------------------- solver */
struct rqb_solver {
  int K, Kparams, dev;
  size_t T, pitch;
  rqb_params P;
  uint32_t max_in, max_out;       /* rows this owner asked for                      */
  uint32_t in_cap, out_cap; /* rows the pinned buffers really have (recycled contexts) */
  int busy;   /* work has been queued on the stream since the last wait */
  int broken; /* a device call failed: do not recycle                    */
  int pages_pending;             /* the pinned page staging is still being copied */
  struct rqb_solver *next_shell;
  void *stream, *ev0, *ev1, *ev2, *ev3;
  /* ordering between streams for batched launches (no timing): ev_ready is recorded on this
   * solver's stream when its uploads must precede a launch on another stream, ev_done on the
   * launching stream after a batched kernel that works on other solvers' rows */
  void *ev_ready, *ev_done;
  struct enc_plan *enc; /* the cached encoder program attached to this context (reference counted) */
  int flavour;          /* RQB_FLAVOUR_*: which flavour of the program to ask the planner for */
  uint8_t *h_in, *h_sym; /* pinned: staging rows of the input space, mirror of the emitted symbols */
  uint32_t *h_flag;      /* pinned word the stream sets when it has drained (rqb_stream_wait_flag) */
  uint32_t flag_seq;
  /* every row of the block lives in ONE device arena (rqb_program.h):
   * [IN: max_in | SYM: max_out | C: L | ZERO: 1 | WS: working rows, grown on demand] */
  uint8_t *d_arena;
  size_t arena_cap; /* bytes */
  uint32_t row0[4], zero_row;
  uint32_t *h_isi; /* pinned: the ISIs of rqb_solver_emit, read in place by the LT kernel */
  int isi_pending; /* an LT kernel reading h_isi has been queued since the last wait */
  uint8_t *d_bounce;            /* rows in the caller's pitch on their way to / from the arena's pitch */
  size_t bounce_cap;
  uint32_t *h_pairs, pairs_cap; /* pinned: (input row, emitted row) pairs of rqb_solver_copy_in_to_sym */
  int pairs_pending;            /* a copy kernel reading h_pairs has been queued since the last wait */
  int batch_args_pending;       /* a batched launch's copy of h_args[1..] has been queued since the last wait */
  /* current program */
  rqb_plan *plan; /* owned unless shared */
  int plan_shared, has_c, timed;
  int want_timing;       /* record CUDA events around the solve launch (rqb_solver_set_timing) */
  int args_valid;        /* d_args holds h_args[0] */
  uint32_t zero_row_set; /* arena row cleared as the ZERO row, +1 (0 = none yet) */
  uint8_t *d_pages, *h_pages; /* device copy / pinned staging of the program pages */
  size_t d_pages_cap, h_pages_cap;
  const uint8_t *cur_pages; /* device pointer actually used (own or cached) */
  rqb_solve_args *h_args, *d_args; /* pinned / device, one entry (batch uses [n]) */
  uint32_t n_out_last;
};

size_t rqb_solver_pitch(const rqb_solver *s) { return s->pitch; }
uint8_t *rqb_solver_staging(rqb_solver *s) { return s->h_in; }
const uint8_t *rqb_solver_sym_mirror(rqb_solver *s) { return s->h_sym; }

/* Solver contexts are recycled whole (stream, events, pinned and device buffers):
 * creating and destroying CUDA objects takes the driver's global lock, and with a
 * dozen decoder threads coming and going at wire rate that lock was where they
 * met (0.6 ms per block at 16 threads, 6 ms at 32).  A destroyed solver goes to a
 * free list and is handed out again to the next create with the same device and
 * symbol size whose rows fit. */
static int solver_wait(rqb_solver *s);
static rqb_solver *g_shells, **g_shells_tail = &g_shells;

/* arena rows a fresh context gets: the fixed spaces plus room for the working rows of
 * a typical program (about 3.5 L, up to 6 L with the table rows of the back-substitution; 8 L reserved; the arena is regrown if a program needs more) */
static long g_ws_reserve = -1; /* NANORQ_B200_WS_RESERVE (tests: start small to exercise the regrowth); -1 = not set */
static pthread_once_t g_ws_once = PTHREAD_ONCE_INIT;
static void ws_reserve_init(void) {
  const char *e = getenv("NANORQ_B200_WS_RESERVE");
  if (e && *e) g_ws_reserve = (long)strtoul(e, NULL, 10);
}
static size_t arena_rows_wanted(uint32_t max_in, uint32_t max_out, const rqb_params *P) {
  size_t ws = 8 * (size_t)P->L + 1024;
  pthread_once(&g_ws_once, ws_reserve_init);
  if (g_ws_reserve >= 0) ws = (size_t)g_ws_reserve;
  return (size_t)max_in + max_out + (size_t)P->L + 1 + ws;
}

/* place the row spaces for this owner's max_in / max_out and clear the ZERO row */
static int solver_layout(rqb_solver *s) {
  s->row0[RQB_SP_IN] = 0;
  s->row0[RQB_SP_SYM] = s->max_in;
  s->row0[RQB_SP_C] = s->max_in + s->max_out;
  s->zero_row = s->row0[RQB_SP_C] + (uint32_t)s->P.L;
  s->row0[RQB_SP_WS] = s->zero_row + 1;
  if (s->zero_row_set == s->zero_row + 1) return 0; /* same layout as this context's last owner: still zero */
  s->zero_row_set = s->zero_row + 1;
  s->busy = 1;
  return rqb_dev_memset(s->d_arena + (size_t)s->zero_row * s->pitch, 0, s->pitch, s->stream);
}
#define ROW_PTR(s, space, k) ((s)->d_arena + ((size_t)(s)->row0[space] + (k)) * (s)->pitch)
static pthread_mutex_t g_shell_mu = PTHREAD_MUTEX_INITIALIZER;

static size_t solver_bytes(const rqb_solver *s) { /* pinned + device memory a context holds */
  return (size_t)s->in_cap * s->pitch + s->arena_cap + (size_t)s->out_cap * s->pitch + s->d_pages_cap +
         s->h_pages_cap + (size_t)s->out_cap * 4 + (size_t)s->pairs_cap * 8 + s->bounce_cap + RQB_ARGS_BYTES;
}

static void solver_detach_plan(rqb_solver *s) {
  if (s->plan && !s->plan_shared) rqb_plan_free(s->plan);
  s->plan = NULL;
  s->plan_shared = 0;
  if (s->enc) enc_plan_unref(s->enc);
  s->enc = NULL;
  s->cur_pages = NULL;
}

/* give everything back to the driver (after the stream has drained) */
static void solver_release(rqb_solver *s) {
  int cur = rqb_dev_get();
  if (cur != s->dev) rqb_dev_set(s->dev);
  if (s->stream) rqb_stream_sync(s->stream);
  solver_detach_plan(s);
  buf_free(s->h_in, 1);
  buf_free(s->d_arena, 0);
  buf_free(s->h_sym, 1);
  buf_free(s->h_flag, 1);
  buf_free(s->h_pairs, 1);
  buf_free(s->d_bounce, 0);
  buf_free(s->h_isi, 1);
  buf_free(s->d_pages, 0);
  buf_free(s->h_pages, 1);
  free(s->h_args);
  buf_free(s->d_args, 0);
  if (s->ev0) rqb_event_destroy(s->ev0);
  if (s->ev1) rqb_event_destroy(s->ev1);
  if (s->ev2) rqb_event_destroy(s->ev2);
  if (s->ev3) rqb_event_destroy(s->ev3);
  if (s->ev_ready) rqb_event_destroy(s->ev_ready);
  if (s->ev_done) rqb_event_destroy(s->ev_done);
  if (s->stream) rqb_stream_destroy(s->stream);
  if (cur != s->dev && cur >= 0) rqb_dev_set(cur);
  free(s);
}

/* Parks the context for the next create of the same shape (it does not wait: work may still be
 * queued on the stream -- an encoder whose repair symbols were never asked for -- every buffer
 * that work touches belongs to the context, and whoever takes it out of the list waits first).
 * Parked contexts count against the cache limit; the oldest ones beyond it are really freed. */
void rqb_solver_destroy(rqb_solver *s) {
  if (!s) return;
  if (s->plan && !s->plan_shared) { /* an encoder program stays attached (and referenced) while queued work may read it */
    rqb_plan_free(s->plan);
    s->plan = NULL;
  }
  if (s->broken) {
    solver_release(s);
    return;
  }
  s->next_shell = NULL;
  rqb_solver *evict = NULL, **evict_tail = &evict;
  pthread_mutex_lock(&g_shell_mu);
  if (s->busy || !g_shells) { /* work still queued: at the tail (oldest first) */
    *g_shells_tail = s;
    g_shells_tail = &s->next_shell;
  } else { /* idle: at the head (most recently used first) */
    s->next_shell = g_shells;
    g_shells = s;
  }
  g_shell_bytes += solver_bytes(s);
  while (g_pool_bytes + g_shell_bytes > cache_limit()) {
    /* over the limit: the least recently used idle context goes (the last idle one in the list), or,
     * when none is idle, the oldest one with work still queued; the one just parked only when nothing
     * else is left (it alone is larger than the limit) */
    rqb_solver **victim = NULL, **self = NULL;
    for (rqb_solver **pp = &g_shells; *pp; pp = &(*pp)->next_shell) {
      if (*pp == s) {
        self = pp;
        continue;
      }
      if (!(*pp)->busy) victim = pp;
      else if (!victim) victim = pp;
    }
    if (!victim) victim = self;
    if (!victim) break;
    rqb_solver *old = *victim;
    *victim = old->next_shell;
    if (g_shells_tail == &old->next_shell) g_shells_tail = victim;
    g_shell_bytes -= solver_bytes(old);
    old->next_shell = NULL;
    *evict_tail = old;
    evict_tail = &old->next_shell;
  }
  pthread_mutex_unlock(&g_shell_mu);
  while (evict) {
    rqb_solver *n = evict->next_shell;
    solver_release(evict);
    evict = n;
  }
}

/* really free every cached solver context, pooled buffer, cached encoder program and recycled plan
 * object (tests, long-lived hosts).  Must not run concurrently with other calls into the library. */
void rqb_release_cached(void) {
  pthread_mutex_lock(&g_shell_mu);
  rqb_solver *s = g_shells;
  g_shells = NULL;
  g_shells_tail = &g_shells;
  pthread_mutex_unlock(&g_shell_mu);
  while (s) {
    rqb_solver *n = s->next_shell;
    g_shell_bytes -= solver_bytes(s);
    solver_release(s);
    s = n;
  }
  pool_drain();
  pthread_mutex_lock(&g_plan_mu);
  enc_plans_trim(0);
  pthread_mutex_unlock(&g_plan_mu);
  rqb_plan_pool_drain();
}

int rqb_solver_create(rqb_solver **out, int K, size_t T, uint32_t max_in, uint32_t max_out) {
  return rqb_solver_create_ex(out, K, K, T, max_in, max_out);
}

int rqb_solver_create_ex(rqb_solver **out, int K, int Kparams, size_t T, uint32_t max_in, uint32_t max_out) {
  return rqb_solver_create_on(out, -1, K, Kparams, T, max_in, max_out);
}

int rqb_solver_device(const rqb_solver *s) { return s->dev; }
void rqb_solver_set_flavour(rqb_solver *s, int flavour) { s->flavour = flavour == RQB_FLAVOUR_HBM ? RQB_FLAVOUR_HBM : RQB_FLAVOUR_AUTO; }

static pthread_once_t g_usolve_hook_once = PTHREAD_ONCE_INIT;
static void usolve_hook_install(void) {
  if (!rqb_plan_usolve_hook) rqb_plan_usolve_hook = usolve_on_device;
}

int rqb_solver_create_on(rqb_solver **out, int want_dev, int K, int Kparams, size_t T, uint32_t max_in,
                         uint32_t max_out) {
  *out = NULL;
  rqb_params P;
  if (K < 1 || rqb_params_init(Kparams, &P) || K > P.Kprime || T == 0 || T > 65535 || max_in < (uint32_t)K) {
    snprintf(g_err, sizeof(g_err), "rqb_solver_create: bad arguments");
    return RQB_E_ARG;
  }
  if (rqb_dev_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device visible: the nanorq_b200 hot path has no CPU fallback");
    return RQB_E_NODEVICE;
  }
  if (!max_out) max_out = 1;
  pthread_once(&g_usolve_hook_once, usolve_hook_install); /* a GPU is present: the planner may use it */
  const int dev = want_dev >= 0 ? want_dev : rqb_dev_default();
  if (dev >= rqb_dev_count()) {
    snprintf(g_err, sizeof(g_err), "rqb_solver_create: no CUDA device %d", dev);
    return RQB_E_ARG;
  }
  rqb_solver *s = NULL;
  pthread_mutex_lock(&g_shell_mu);
  {
    /* never a context much larger than asked for: encoders and decoders of one block size must not
     * keep taking each other's contexts.  Idle contexts are parked at the head of the list (most
     * recently used first: warm, and found at once -- tiny blocks come and go every few microseconds
     * and this walk happens under a process-wide lock); contexts with work still queued are parked
     * at the tail (oldest first: most likely done) and only taken when no idle one fits. */
    rqb_solver **best = NULL;
    const size_t want_bytes = arena_rows_wanted(max_in, max_out, &P);
    for (rqb_solver **pp = &g_shells; *pp; pp = &(*pp)->next_shell) {
      rqb_solver *c = *pp;
      if (c->dev != dev || c->T != T || c->in_cap < max_in || c->out_cap < max_out) continue;
      if (c->arena_cap < want_bytes * c->pitch) continue;
      if ((size_t)c->in_cap > (size_t)max_in + max_in / 4 + 64 || (size_t)c->out_cap > (size_t)max_out + max_out / 4 + 64) continue;
      if (!c->busy) {
        best = pp;
        break;
      }
      if (!best) best = pp;
    }
    if (best) {
      s = *best;
      *best = s->next_shell;
      if (g_shells_tail == &s->next_shell) g_shells_tail = best;
      g_shell_bytes -= solver_bytes(s);
    }
  }
  pthread_mutex_unlock(&g_shell_mu);
  if (s) {
    s->K = K;
    s->Kparams = Kparams;
    s->P = P;
    s->max_in = max_in;
    s->max_out = max_out;
    s->has_c = s->timed = s->want_timing = 0;
    s->flavour = RQB_FLAVOUR_AUTO;
    s->n_out_last = 0;
    s->next_shell = NULL;
    if (bind_dev(s->dev)) { /* the context is already out of the list: do not leak it */
      s->broken = 1;
      rqb_solver_destroy(s);
      return dev_fail(0, "rqb_solver_create bind");
    }
    if (s->busy && solver_wait(s)) { /* the previous owner left work queued */
      s->broken = 1;
      rqb_solver_destroy(s);
      return RQB_E_NODEVICE;
    }
    solver_detach_plan(s); /* nothing queued can read the previous owner's program any more */
    if (solver_layout(s)) {
      s->broken = 1;
      rqb_solver_destroy(s);
      return RQB_E_NODEVICE;
    }
    *out = s;
    return 0;
  }
  g_cnt_ctx_new++;
  s = calloc(1, sizeof(*s));
  if (!s) {
    snprintf(g_err, sizeof(g_err), "rqb_solver_create: out of memory");
    return RQB_E_ARG;
  }
  s->K = K;
  s->Kparams = Kparams;
  s->T = T;
  s->pitch = round_up(T, 64);
  s->P = P;
  s->max_in = s->in_cap = max_in;
  s->max_out = s->out_cap = max_out;
  s->dev = dev;
  s->arena_cap = pool_class(arena_rows_wanted(max_in, max_out, &P) * s->pitch);
  int e = bind_dev(s->dev);
  e = e ? e : rqb_stream_create(&s->stream);
  e = e ? e : rqb_event_create(&s->ev2);
  e = e ? e : rqb_event_create(&s->ev3);
  e = e ? e : rqb_event_create(&s->ev0);
  e = e ? e : rqb_event_create(&s->ev1);
  e = e ? e : rqb_event_create_sync(&s->ev_ready);
  e = e ? e : rqb_event_create_sync(&s->ev_done);
  e = e ? e : pool_get((void **)&s->h_in, (size_t)s->in_cap * s->pitch, 1);
  e = e ? e : pool_get((void **)&s->d_arena, s->arena_cap, 0);
  e = e ? e : pool_get((void **)&s->h_sym, (size_t)s->out_cap * s->pitch, 1);
  e = e ? e : pool_get((void **)&s->h_flag, 64, 1);
  if (!e) *s->h_flag = s->flag_seq = 0;
  e = e ? e : pool_get((void **)&s->h_isi, (size_t)s->out_cap * 4, 1);
  /* small control blocks are sent from PAGEABLE memory on purpose: cudaMemcpyAsync
   * stages a pageable source before it returns, so these buffers may be rewritten
   * for the next launch while earlier copies are still queued on the stream */
  s->h_args = calloc(1, RQB_ARGS_BYTES);
  if (!s->h_args) {
    snprintf(g_err, sizeof(g_err), "rqb_solver_create: out of memory");
    s->broken = 1;
    rqb_solver_destroy(s);
    return RQB_E_ARG;
  }
  e = e ? e : pool_get((void **)&s->d_args, RQB_ARGS_BYTES, 0);
  e = e ? e : solver_layout(s);
  if (e) {
    dev_fail(e, "rqb_solver_create allocation");
    s->broken = 1;
    rqb_solver_destroy(s);
    return RQB_E_NODEVICE;
  }
  /* pad bytes of the staging rows must be zero: they are transformed too (pool
   * buffers are recycled, symbol bytes are always overwritten before use) */
  if (s->pitch > T)
    for (uint32_t r = 0; r < s->in_cap; r++) memset(s->h_in + (size_t)r * s->pitch + T, 0, s->pitch - T);
  *out = s;
  return 0;
}

static int solver_wait(rqb_solver *s) {
  /* a batched launch on another solver's stream that works on this solver's rows is ordered
   * before anything queued here afterwards by ev_done (rqb_solver_run_batch_on) */
  DEV(rqb_stream_wait_flag(s->stream, s->h_flag, &s->flag_seq));
  s->busy = 0;
  s->pages_pending = 0;
  s->pairs_pending = 0;
  s->isi_pending = 0;
  s->batch_args_pending = 0;
  return 0;
}

int rqb_solver_upload(rqb_solver *s, uint32_t first, uint32_t n) {
  BIND(s->dev);
  if ((uint64_t)first + n > s->max_in) return RQB_E_ARG;
  if (!n) return 0;
  rqb_copy_fence(); /* the staging rows may have been written with non-temporal stores */
  s->busy = 1;
  DEV(rqb_copy_h2d(ROW_PTR(s, RQB_SP_IN, first), s->h_in + (size_t)first * s->pitch, (size_t)n * s->pitch,
                   s->stream));
  return 0;
}

static int ensure_cap(rqb_solver *s, void **buf, size_t *cap, size_t need, int pinned);

int rqb_solver_upload_rows(rqb_solver *s, uint32_t first, uint32_t n, const uint8_t *src, size_t src_pitch) {
  BIND(s->dev);
  if ((uint64_t)first + n > s->max_in || !src || src_pitch < s->T) return RQB_E_ARG;
  if (!n) return 0;
  rqb_copy_fence();
  s->busy = 1;
  /* pad bytes of the device rows are never read back and never mix with payload bytes (row operations
   * are column-local), so only the T payload bytes of each row travel */
  /* Host<->device copies are always LINEAR: a 2-D DMA of ~1 KB rows reaches a fifth of the link's
   * bandwidth (measured: 10 GB/s against 53).  Equal pitches: one copy.  Otherwise the rows cross the
   * link as they lie in the caller's memory, into a bounce buffer, and a kernel re-pitches them. */
  const size_t span = (size_t)(n - 1) * src_pitch + s->T;
  if (src_pitch == s->pitch) {
    DEV(rqb_copy_h2d(ROW_PTR(s, RQB_SP_IN, first), src, span, s->stream));
  } else if (n < 16 || span > ((size_t)64 << 20)) {
    DEV(rqb_copy2d_h2d(ROW_PTR(s, RQB_SP_IN, first), s->pitch, src, src_pitch, s->T, n, s->stream));
  } else {
    int rc = ensure_cap(s, (void **)&s->d_bounce, &s->bounce_cap, span, 0);
    if (rc) return rc;
    DEV(rqb_copy_h2d(s->d_bounce, src, span, s->stream));
    DEV(rqb_launch_repitch(ROW_PTR(s, RQB_SP_IN, first), s->pitch, s->d_bounce, src_pitch, (uint32_t)s->T, n, s->stream));
  }
  return 0;
}

int rqb_solver_fetch_rows(rqb_solver *s, int space, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch,
                          int wait) {
  BIND(s->dev);
  const uint32_t cap = space == RQB_SP_IN ? s->max_in : space == RQB_SP_SYM ? s->max_out : (uint32_t)s->P.L;
  if (space < RQB_SP_IN || space > RQB_SP_C || (uint64_t)first + n > cap || !dst || dst_pitch < s->T) return RQB_E_ARG;
  if (space == RQB_SP_C && !s->has_c) return RQB_E_ARG;
  if (n) {
    s->busy = 1;
    const size_t span = (size_t)(n - 1) * dst_pitch + s->T;
    if (dst_pitch == s->pitch) {
      DEV(rqb_copy_d2h(dst, ROW_PTR(s, space, first), span, s->stream));
    } else if (n < 16 || dst_pitch != s->T || span > ((size_t)64 << 20)) {
      /* a destination with gaps between its rows keeps its gaps: 2-D copy */
      DEV(rqb_copy2d_d2h(dst, dst_pitch, ROW_PTR(s, space, first), s->pitch, s->T, n, s->stream));
    } else { /* packed destination: pack on the device, one linear copy */
      int rc = ensure_cap(s, (void **)&s->d_bounce, &s->bounce_cap, span, 0);
      if (rc) return rc;
      DEV(rqb_launch_repitch(s->d_bounce, dst_pitch, ROW_PTR(s, space, first), s->pitch, (uint32_t)s->T, n, s->stream));
      DEV(rqb_copy_d2h(dst, s->d_bounce, span, s->stream));
    }
  }
  if (wait) {
    int w = solver_wait(s);
    if (w) return w;
  }
  return 0;
}

int rqb_solver_copy_in_to_sym(rqb_solver *s, const uint32_t *sym_row, const uint32_t *in_row, uint32_t n) {
  BIND(s->dev);
  if (!n) return 0;
  if (n > s->max_out) return RQB_E_ARG;
  /* The pairs live in page-locked host memory that the kernel reads in place (unified addressing):
   * a copy from pageable memory would make the driver wait for everything queued on the stream --
   * here the solve itself -- before it returns, and it does that under a lock other threads'
   * launches need (measured: 0.6 ms per block, serialised over all threads). */
  if (s->pairs_cap < n) {
    if (s->h_pairs) {
      int w = solver_wait(s);
      if (w) return w;
      pool_put(s->h_pairs, (size_t)s->pairs_cap * 8, 1);
      s->h_pairs = NULL;
      s->pairs_cap = 0;
    }
    DEV(pool_get((void **)&s->h_pairs, (size_t)s->max_out * 8, 1));
    s->pairs_cap = s->max_out;
  } else if (s->pairs_pending) { /* an earlier copy kernel may still be reading the buffer */
    int w = solver_wait(s);
    if (w) return w;
  }
  for (uint32_t k = 0; k < n; k++) {
    if (in_row[k] >= s->max_in || sym_row[k] >= s->max_out) return RQB_E_ARG;
    s->h_pairs[2 * k] = s->row0[RQB_SP_IN] + in_row[k];
    s->h_pairs[2 * k + 1] = s->row0[RQB_SP_SYM] + sym_row[k];
  }
  s->busy = 1;
  s->pairs_pending = 1;
  DEV(rqb_launch_copy_rows(s->d_arena, s->pitch, s->h_pairs, n, (uint32_t)round_up(s->T, 16), s->stream));
  return 0;
}

void *rqb_host_alloc(size_t bytes) {
  void *p = NULL;
  if (rqb_dev_count() <= 0 || rqb_host_malloc(&p, bytes)) return NULL;
  return p;
}
void rqb_host_release(void *p) {
  if (p) rqb_host_free(p);
}
int rqb_host_pin(void *p, size_t bytes) {
  if (rqb_dev_count() <= 0) return RQB_E_NODEVICE;
  DEV(rqb_host_register(p, bytes));
  return 0;
}
int rqb_host_unpin(void *p) {
  DEV(rqb_host_unregister(p));
  return 0;
}

static void fill_stats(const rqb_plan *p, rqb_solver_stats *o) {
  memset(o, 0, sizeof(*o));
  o->i = p->st.i; o->u = p->st.u; o->nb = p->st.nb; o->rho = p->st.rho; o->nfree = p->st.nfree;
  o->levels_fwd = p->st.levels_fwd; o->n_levels = p->st.n_levels; o->n_tasks = p->st.n_tasks;
  o->n_pages = p->st.n_pages; o->n_srcs = p->st.n_srcs; o->n_gf_srcs = p->st.n_gf_srcs;
  o->n_horner = p->st.n_horner; o->nnz = p->st.nnz; o->t_matrix = p->st.t_matrix;
  o->t_peel = p->st.t_peel; o->t_dense = p->st.t_dense; o->t_emit = p->st.t_emit;
  o->n_ws_rows = p->n_ws_rows;
  o->n_parts = p->st.n_parts;
  o->slice_bytes = p->smem ? (int)p->slice_bytes : (int)RQB_SLICE_BYTES;
  o->smem = p->smem;
  o->n_slots = p->n_slots;
  o->tab_bits = p->tab_bits;
}

int rqb_solver_get_stats(const rqb_solver *s, rqb_solver_stats *out) {
  if (!s->plan) return RQB_E_ARG;
  fill_stats(s->plan, out);
  return 0;
}

/* make room for the program's working rows; buffers are only handed back to the
 * pool once nothing queued on the stream can still touch them */
static int ensure_cap(rqb_solver *s, void **buf, size_t *cap, size_t need, int pinned) {
  if (need <= *cap) return 0;
  if (*buf) {
    { int _w = solver_wait(s); if (_w) return _w; }
    pool_put(*buf, *cap, pinned);
    *buf = NULL;
    *cap = 0;
  }
  need += need / 4; /* programs of one block size differ a little from block to block: do not regrow for each */
  DEV(pool_get(buf, need, pinned));
  *cap = pool_class(need);
  return 0;
}

static int solver_set_args(rqb_solver *s) {
  const rqb_plan *p = s->plan;
  if (memcmp(p->row0, s->row0, sizeof(s->row0)) || p->zero_row != s->zero_row) {
    snprintf(g_err, sizeof(g_err), "solve program was built for another arena layout");
    return RQB_E_ARG;
  }
  size_t need = (size_t)p->n_rows * s->pitch;
  if (need > s->arena_cap) {
    /* the program needs more working rows than the arena has: move the fixed spaces
     * (uploaded symbols included) into a larger one */
    uint8_t *bigger = NULL;
    size_t cap = pool_class(need + need / 4);
    g_cnt_regrow++;
    int w = solver_wait(s);
    if (w) return w;
    DEV(pool_get((void **)&bigger, cap, 0));
    DEV(rqb_copy_d2d(bigger, s->d_arena, (size_t)s->row0[RQB_SP_WS] * s->pitch, s->stream));
    w = solver_wait(s);
    if (w) return w;
    pool_put(s->d_arena, s->arena_cap, 0);
    s->d_arena = bigger;
    s->arena_cap = cap;
    /* the ZERO row sits below the working rows and was copied with the fixed spaces */
  }
  rqb_solve_args *a = s->h_args, na;
  memset(&na, 0, sizeof(na));
  na.base = s->d_arena;
  na.pages = s->cur_pages;
  na.pitch = (uint32_t)s->pitch;
  na.n_pages = p->n_pages;
  na.width = (uint32_t)round_up(s->T, 16);
  na.pad = p->smem ? p->n_slots : 0; /* shared-memory flavour: slots the CTA needs */
  if (!s->args_valid || memcmp(a, &na, sizeof(na))) { /* a recycled encoder context usually has them on the device already */
    *a = na;
    s->busy = 1;
    DEV(rqb_copy_h2d(s->d_args, s->h_args, sizeof(*a), s->stream));
    s->args_valid = 1;
  }
  s->has_c = p->n_c_rows != 0;
  s->n_out_last = p->n_out;
  return 0;
}

int rqb_solver_plan(rqb_solver *s, const rqb_solve_request *req) {
  BIND(s->dev);
  if (req->n_out > s->max_out) return RQB_E_ARG;
  rqb_plan_request pr;
  pr.K = s->Kparams;
  pr.overhead = req->overhead;
  pr.isi = req->isi;
  pr.in_row = req->in_row;
  pr.want_c = req->want_c;
  pr.n_out = (int)req->n_out;
  pr.out_isi = req->out_isi;
  pr.in_rows = s->max_in;
  pr.sym_rows = s->max_out;
  /* the planner writes the program straight into this solver's pinned page buffer (no copy of
   * ~1 MB per block); the buffer must not still be the source of the previous program's upload */
  {
    size_t want = (size_t)400 * (size_t)s->P.L; /* measured: ~300 bytes of program per intermediate symbol */
    if (s->h_pages_cap < want || s->d_pages_cap < want) {
      int rc0 = ensure_cap(s, (void **)&s->d_pages, &s->d_pages_cap, want, 0);
      rc0 = rc0 ? rc0 : ensure_cap(s, (void **)&s->h_pages, &s->h_pages_cap, want, 1);
      if (rc0) return rc0;
    }
    if (s->pages_pending) {
      int w = solver_wait(s);
      if (w) return w;
    }
  }
  pr.out_row = req->out_row;
  pr.smem_budget = s->flavour == RQB_FLAVOUR_HBM ? 0u : solver_smem_budget();
  pr.pages_buf = s->h_pages;
  pr.pages_buf_cap = s->h_pages_cap < s->d_pages_cap ? s->h_pages_cap : s->d_pages_cap;
  rqb_plan *p = NULL;
  PF_T0;
  int rc = rqb_plan_build(&pr, &p);
  PF(RQB_PF_REP_PLAN);
  if (rc == 1) return RQB_NEED_MORE;
  if (rc) {
    snprintf(g_err, sizeof(g_err), "rqb_plan_build failed (%d)", rc);
    return rc == -4 ? RQB_E_TOOBIG : RQB_E_ARG;
  }
  if (s->enc) { /* queued work may still read the cached encoder program */
    int w = solver_wait(s);
    if (w) {
      rqb_plan_free(p);
      return w;
    }
  }
  solver_detach_plan(s);
  s->plan = p;
  s->plan_shared = 0;
  size_t pb = (size_t)p->n_pages * RQB_PAGE_BYTES;
  if (p->pages != s->h_pages) { /* the program did not fit the pinned buffer: grow it and copy */
    size_t want = pb + pb / 4;
    rc = ensure_cap(s, (void **)&s->d_pages, &s->d_pages_cap, want, 0);
    rc = rc ? rc : ensure_cap(s, (void **)&s->h_pages, &s->h_pages_cap, want, 1);
    if (rc) return rc;
    memcpy(s->h_pages, p->pages, pb);
  }
  s->pages_pending = 1;
  s->busy = 1;
  DEV(rqb_copy_h2d(s->d_pages, s->h_pages, pb, s->stream));
  s->cur_pages = s->d_pages;
  PF(RQB_PF_REP_PAGES);
  rc = solver_set_args(s);
  PF(RQB_PF_REP_ARGS);
  return rc;
}

/* ---------------------------------------------------- planning several blocks */
typedef struct {
  rqb_solver **sv;
  const rqb_solve_request *reqs;
  int n, *rc;
  _Atomic int next, good;
} plan_batch;

static void *plan_batch_worker(void *arg) {
  plan_batch *pb = arg;
  for (;;) {
    const int k = pb->next++;
    if (k >= pb->n) break;
    const int rc = rqb_solver_plan(pb->sv[k], &pb->reqs[k]);
    if (pb->rc) pb->rc[k] = rc;
    if (rc == 0) pb->good++;
  }
  return NULL;
}

/* Worker threads are kept (the planner's per-thread scratch arenas and the u-solve contexts stay warm:
 * a fresh thread would allocate and fault in megabytes for every call).  One batch at a time uses the
 * pool; a caller that finds it taken plans in its own thread. */
#define PLAN_POOL_MAX 64
static struct {
  pthread_t th;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  plan_batch *job;
} g_pw[PLAN_POOL_MAX];
static int g_npw;
static pthread_mutex_t g_pool_use = PTHREAD_MUTEX_INITIALIZER, g_done_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_done_cv = PTHREAD_COND_INITIALIZER;
static int g_pending;

static void *plan_pool_worker(void *arg) {
  const int me = (int)(intptr_t)arg;
  for (;;) {
    pthread_mutex_lock(&g_pw[me].mu);
    while (!g_pw[me].job) pthread_cond_wait(&g_pw[me].cv, &g_pw[me].mu);
    plan_batch *job = g_pw[me].job;
    g_pw[me].job = NULL;
    pthread_mutex_unlock(&g_pw[me].mu);
    plan_batch_worker(job);
    pthread_mutex_lock(&g_done_mu);
    if (--g_pending == 0) pthread_cond_signal(&g_done_cv);
    pthread_mutex_unlock(&g_done_mu);
  }
  return NULL;
}

int rqb_solver_plan_batch(rqb_solver **solvers, const rqb_solve_request *reqs, int n, int nthreads, int *rc) {
  if (n <= 0 || !solvers || !reqs) return 0;
  plan_batch pb = {solvers, reqs, n, rc, 0, 0};
  if (nthreads > n) nthreads = n;
  if (nthreads > PLAN_POOL_MAX + 1) nthreads = PLAN_POOL_MAX + 1;
  if (nthreads > 1 && pthread_mutex_trylock(&g_pool_use) == 0) {
    while (g_npw < nthreads - 1) { /* the calling thread is worker 0 */
      pthread_mutex_init(&g_pw[g_npw].mu, NULL);
      pthread_cond_init(&g_pw[g_npw].cv, NULL);
      g_pw[g_npw].job = NULL;
      if (pthread_create(&g_pw[g_npw].th, NULL, plan_pool_worker, (void *)(intptr_t)g_npw) != 0) break;
      pthread_detach(g_pw[g_npw].th);
      g_npw++;
    }
    const int helpers = g_npw < nthreads - 1 ? g_npw : nthreads - 1;
    pthread_mutex_lock(&g_done_mu);
    g_pending = helpers;
    pthread_mutex_unlock(&g_done_mu);
    for (int t = 0; t < helpers; t++) {
      pthread_mutex_lock(&g_pw[t].mu);
      g_pw[t].job = &pb;
      pthread_cond_signal(&g_pw[t].cv);
      pthread_mutex_unlock(&g_pw[t].mu);
    }
    plan_batch_worker(&pb);
    pthread_mutex_lock(&g_done_mu);
    while (g_pending > 0) pthread_cond_wait(&g_done_cv, &g_done_mu);
    pthread_mutex_unlock(&g_done_mu);
    pthread_mutex_unlock(&g_pool_use);
  } else {
    plan_batch_worker(&pb);
  }
  return pb.good;
}

static _Atomic int g_plan_threads = 1;
void rqb_set_plan_threads(int n) { g_plan_threads = n < 1 ? 1 : n; }
int rqb_get_plan_threads(void) { return g_plan_threads; }

int rqb_solver_plan_encode(rqb_solver *s, int want_c, uint32_t n_rep) {
  BIND(s->dev);
  if (n_rep > s->max_out) return RQB_E_ARG;
  pthread_mutex_lock(&g_plan_mu);
  enc_plan *e = g_enc_plans;
  for (; e; e = e->next)
    if (e->K == s->K && e->Kparams == s->Kparams && e->want_c == want_c && e->n_out == n_rep && e->dev == s->dev &&
        e->in_rows == s->max_in && e->sym_rows == s->max_out && e->hbm == (s->flavour == RQB_FLAVOUR_HBM))
      break;
  if (!e) {
    const int Kp = s->P.Kprime;
    uint32_t *isi = malloc(sizeof(uint32_t) * (size_t)Kp), *in_row = malloc(sizeof(uint32_t) * (size_t)Kp);
    uint32_t *oi = malloc(sizeof(uint32_t) * (size_t)(n_rep ? n_rep : 1));
    rqb_plan *p = NULL;
    int rc = -7;
    if (isi && in_row && oi) {
      for (int k = 0; k < Kp; k++) {
        isi[k] = (uint32_t)k;
        in_row[k] = k < s->K ? (uint32_t)k : RQB_ROW_NONE;
      }
      for (uint32_t k = 0; k < n_rep; k++) oi[k] = (uint32_t)Kp + k; /* repair ESI K+k <-> ISI K'+k */
      rqb_plan_request pr = {s->Kparams, 0, isi, in_row, want_c, (int)n_rep, oi, s->max_in, s->max_out, NULL, 0,
                             NULL, s->flavour == RQB_FLAVOUR_HBM ? 0u : solver_smem_budget()};
      rc = rqb_plan_build(&pr, &p);
    }
    free(isi);
    free(in_row);
    free(oi);
    if (rc) {
      pthread_mutex_unlock(&g_plan_mu);
      snprintf(g_err, sizeof(g_err), "encoder plan for K=%d failed (%d)", s->K, rc);
      return rc == 1 ? RQB_NEED_MORE : (rc == -4 ? RQB_E_TOOBIG : RQB_E_ARG);
    }
    e = calloc(1, sizeof(*e));
    size_t pb = (size_t)p->n_pages * RQB_PAGE_BYTES;
    int de = e ? 0 : -1;
    if (e) {
      e->K = s->K;
      e->Kparams = s->Kparams;
      e->want_c = want_c;
      e->n_out = n_rep;
      e->in_rows = s->max_in;
      e->sym_rows = s->max_out;
      e->dev = s->dev;
      e->hbm = s->flavour == RQB_FLAVOUR_HBM;
      e->plan = p;
      de = rqb_dev_malloc((void **)&e->d_pages, pb);
      de = de ? de : rqb_copy_h2d(e->d_pages, p->pages, pb, s->stream);
      de = de ? de : rqb_stream_sync(s->stream);
    }
    if (de) { /* nothing of the half-built entry is kept */
      if (e && e->d_pages) rqb_dev_free(e->d_pages);
      free(e);
      rqb_plan_free(p);
      pthread_mutex_unlock(&g_plan_mu);
      return dev_fail(de, "encoder plan upload");
    }
    e->next = g_enc_plans;
    g_enc_plans = e;
  }
  e->refs++; /* referenced before the cache is trimmed: the new entry is not a candidate */
  e->use = ++g_enc_clock;
  enc_plans_trim(RQB_ENC_PLANS_MAX);
  pthread_mutex_unlock(&g_plan_mu);
  if (s->enc != e && s->enc && s->busy) { /* queued work may still read the previous program */
    int w = solver_wait(s);
    if (w) {
      enc_plan_unref(e);
      return w;
    }
  }
  solver_detach_plan(s);
  s->enc = e;
  s->plan = e->plan;
  s->plan_shared = 1;
  s->cur_pages = e->d_pages;
  return solver_set_args(s);
}

int rqb_solver_run(rqb_solver *s) {
  if (!s->plan) return RQB_E_ARG;
  BIND(s->dev);
  s->busy = 1;
  if (s->want_timing) DEV(rqb_event_record(s->ev0, s->stream));
  if (s->plan->smem)
    DEV(rqb_launch_solve_smem(s->d_args, 1, s->h_args->width, s->plan->slice_bytes, s->plan->n_slots, s->stream));
  else
    DEV(rqb_launch_solve(s->d_args, 1, s->h_args->width, s->stream));
  if (s->want_timing) DEV(rqb_event_record(s->ev1, s->stream));
  s->timed = s->want_timing;
  return 0;
}

/* CUDA events around every solve launch of this solver (for rqb_solver_last_kernel_ms);
 * off by default: two more driver calls per block matter at wire rate */
void rqb_solver_set_timing(rqb_solver *s, int on) { s->want_timing = on != 0; }

/* time a region of work queued on this solver's stream with CUDA events */
int rqb_solver_mark(rqb_solver *s, int end) {
  BIND(s->dev);
  s->busy = 1;
  DEV(rqb_event_record(end ? s->ev3 : s->ev2, s->stream));
  return 0;
}
int rqb_solver_marked_ms(rqb_solver *s, float *ms) {
  BIND(s->dev);
  DEV(rqb_event_sync(s->ev3));
  DEV(rqb_event_elapsed_ms(s->ev2, s->ev3, ms));
  return 0;
}

/* one launch for n blocks on the stream of `own` (which also lends its argument
 * buffers and timing events) */
int rqb_solver_run_batch_on(rqb_solver **sv, int n, rqb_solver *own) {
  if (n <= 0 || !own) return RQB_E_ARG;
  BIND(own->dev);
  if ((size_t)(n + 1) * sizeof(rqb_solve_args) > RQB_ARGS_BYTES) return RQB_E_ARG;
  /* entry 0 of the owner's buffer stays its own single-run argument block.  The blocks are
   * grouped by the flavour of their program (HBM, or shared memory with 16/32/64-byte slots):
   * one launch per flavour present, normally one. */
  rqb_solve_args *h = own->h_args + 1, *d = own->d_args + 1;
  int cls_n[4] = {0, 0, 0, 0}, cls_at[4];
  uint32_t cls_slots[4] = {0, 0, 0, 0};
#define FLAVOUR_CLASS(p) (!(p)->smem ? 0 : (p)->slice_bytes == 16 ? 1 : (p)->slice_bytes == 32 ? 2 : 3)
  for (int k = 0; k < n; k++) {
    if (!sv[k]->plan || sv[k]->T != own->T || sv[k]->dev != own->dev) return RQB_E_ARG;
    cls_n[FLAVOUR_CLASS(sv[k]->plan)]++;
  }
  cls_at[0] = 0;
  for (int c = 1; c < 4; c++) cls_at[c] = cls_at[c - 1] + cls_n[c - 1];
  if (own->batch_args_pending) {
    /* the copy engine may not have read the argument blocks of the owner's previous batch yet: they
     * are only rewritten when this batch differs from it (a loop over the same blocks does not) */
    int at[4] = {cls_at[0], cls_at[1], cls_at[2], cls_at[3]}, same = 1;
    for (int k = 0; k < n && same; k++) same = !memcmp(&h[at[FLAVOUR_CLASS(sv[k]->plan)]++], sv[k]->h_args, sizeof(*h));
    if (!same) {
      int w = solver_wait(own);
      if (w) return w;
    }
  }
  for (int k = 0; k < n; k++) {
    if (sv[k] != own && sv[k]->busy) {
      /* its uploads (symbols, program pages, arguments) are queued on its own stream: the launching
       * stream waits for them on the device, the host does not */
      DEV(rqb_event_record(sv[k]->ev_ready, sv[k]->stream));
      DEV(rqb_stream_wait_event(own->stream, sv[k]->ev_ready));
    }
    const int c = FLAVOUR_CLASS(sv[k]->plan);
    h[cls_at[c]++] = *sv[k]->h_args;
    if (sv[k]->plan->smem && sv[k]->plan->n_slots > cls_slots[c]) cls_slots[c] = sv[k]->plan->n_slots;
  }
#undef FLAVOUR_CLASS
  own->busy = 1;
  own->batch_args_pending = 1;
  DEV(rqb_copy_h2d(d, h, (size_t)n * sizeof(rqb_solve_args), own->stream));
  if (own->want_timing) DEV(rqb_event_record(own->ev0, own->stream));
  for (int c = 0, at = 0; c < 4; at += cls_n[c], c++) {
    if (!cls_n[c]) continue;
    if (c == 0)
      DEV(rqb_launch_solve(d + at, cls_n[c], h[at].width, own->stream));
    else
      DEV(rqb_launch_solve_smem(d + at, cls_n[c], h[at].width, 8u << c, cls_slots[c], own->stream));
  }
  if (own->want_timing) DEV(rqb_event_record(own->ev1, own->stream));
  own->timed = own->want_timing;
  /* whatever a member queues on its own stream from now on (fetches, emits, uploads of the next
   * symbols, a later launch) runs after this kernel: no entry point needs to know about the batch */
  DEV(rqb_event_record(own->ev_done, own->stream));
  for (int k = 0; k < n; k++)
    if (sv[k] != own) {
      DEV(rqb_stream_wait_event(sv[k]->stream, own->ev_done));
      sv[k]->busy = 1;
    }
  return 0;
}

int rqb_batch_slice_bytes(int nblocks, size_t T) { return rqb_solve_slice_bytes(nblocks, (uint32_t)round_up(T, 16)); }

int rqb_solver_run_batch(rqb_solver **sv, int n) {
  if (n <= 0) return RQB_E_ARG;
  if (n == 1) return rqb_solver_run(sv[0]);
  return rqb_solver_run_batch_on(sv, n, sv[0]);
}

int rqb_solver_emit(rqb_solver *s, const uint32_t *isi, uint32_t n) {
  BIND(s->dev);
  if (!s->has_c || n > s->max_out) return RQB_E_ARG;
  /* the LT kernel reads the ISIs from page-locked host memory in place: no copy from pageable
   * memory, which would wait for the whole stream under the driver's lock (see copy_in_to_sym) */
  if (s->isi_pending) {
    int w = solver_wait(s);
    if (w) return w;
  }
  memcpy(s->h_isi, isi, (size_t)n * 4);
  s->busy = 1;
  s->isi_pending = 1;
  DEV(rqb_launch_lt(&s->P, ROW_PTR(s, RQB_SP_C, 0), (uint32_t)s->pitch, s->h_isi, n, ROW_PTR(s, RQB_SP_SYM, 0),
                    (uint32_t)s->pitch, (uint32_t)round_up(s->T, 16), s->stream));
  s->n_out_last = n;
  return 0;
}

int rqb_solver_sync(rqb_solver *s) {
  BIND(s->dev);
  { int _w = solver_wait(s); if (_w) return _w; }
  return 0;
}

int rqb_solver_last_kernel_ms(rqb_solver *s, float *ms) {
  BIND(s->dev);
  if (!s->timed) return RQB_E_ARG;
  DEV(rqb_event_sync(s->ev1));
  DEV(rqb_event_elapsed_ms(s->ev0, s->ev1, ms));
  return 0;
}

int rqb_solver_fetch_syms(rqb_solver *s, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch) {
  BIND(s->dev);
  if ((uint64_t)first + n > s->max_out) return RQB_E_ARG;
  if (!n) return 0;
  s->busy = 1;
  DEV(rqb_copy_d2h(s->h_sym + (size_t)first * s->pitch, ROW_PTR(s, RQB_SP_SYM, first), (size_t)n * s->pitch,
                   s->stream));
  { int _w = solver_wait(s); if (_w) return _w; }
  if (dst)
    for (uint32_t k = 0; k < n; k++)
      memcpy(dst + (size_t)k * dst_pitch, s->h_sym + (size_t)(first + k) * s->pitch, s->T);
  return 0;
}

/* queue the copy of emitted symbols [first, first+n) into the pinned mirror without
 * waiting; rqb_solver_sync() (or any fetch) completes it */
int rqb_solver_fetch_syms_async(rqb_solver *s, uint32_t first, uint32_t n) {
  BIND(s->dev);
  if ((uint64_t)first + n > s->max_out) return RQB_E_ARG;
  if (!n) return 0;
  s->busy = 1;
  DEV(rqb_copy_d2h(s->h_sym + (size_t)first * s->pitch, ROW_PTR(s, RQB_SP_SYM, first), (size_t)n * s->pitch,
                   s->stream));
  return 0;
}

int rqb_solver_fetch_c(rqb_solver *s, uint32_t first, uint32_t n, uint8_t *dst, size_t dst_pitch) {
  BIND(s->dev);
  if (!s->has_c || (uint64_t)first + n > (uint32_t)s->P.L) return RQB_E_ARG;
  s->busy = 1;
  DEV(rqb_copy2d_d2h(dst, dst_pitch, ROW_PTR(s, RQB_SP_C, first), s->pitch, s->T, n, s->stream));
  { int _w = solver_wait(s); if (_w) return _w; }
  return 0;
}

/* ------------------------------------------------ host-only plan access */
uint32_t rqb_smem_budget(void) { return RQB_SMEM_BUDGET_BYTES; }

int rqb_plan_blob_build(int K, const rqb_solve_request *req, rqb_plan_blob *out) {
  return rqb_plan_blob_build_ex(K, req, 0, out);
}

int rqb_plan_blob_build_ex(int K, const rqb_solve_request *req, uint32_t smem_budget, rqb_plan_blob *out) {
  rqb_params P;
  if (!rqb_plan_usolve_hook && rqb_plan_usolve_mode == 2 && rqb_dev_count() > 0) rqb_plan_usolve_hook = usolve_on_device;
  memset(out, 0, sizeof(*out));
  if (rqb_params_init(K, &P) || req->overhead < 0) return RQB_E_ARG;
  uint32_t in_rows = 1; /* the smallest input space that holds every row the request names */
  for (int k = 0; k < P.Kprime + req->overhead; k++)
    if (req->in_row[k] != RQB_NO_ROW && req->in_row[k] >= in_rows) in_rows = req->in_row[k] + 1;
  rqb_plan_request pr = {K, req->overhead, req->isi, req->in_row, req->want_c, (int)req->n_out, req->out_isi,
                         in_rows, req->n_out ? req->n_out : 1, NULL, 0, req->out_row, smem_budget};
  rqb_plan *p = NULL;
  int rc = rqb_plan_build(&pr, &p);
  if (rc == 1) return RQB_NEED_MORE;
  if (rc) return rc == -4 ? RQB_E_TOOBIG : RQB_E_ARG;
  out->smem = p->smem;
  out->slice_bytes = p->slice_bytes;
  out->n_slots = p->n_slots;
  out->tab_bits = p->tab_bits;
  out->n_ws_rows = p->n_ws_rows;
  memcpy(out->row0, p->row0, sizeof(out->row0));
  out->zero_row = p->zero_row;
  out->n_rows = p->n_rows;
  out->n_pages = p->n_pages;
  out->page_bytes = RQB_PAGE_BYTES;
  out->pages = p->pages;
  out->opaque = p;
  fill_stats(p, &out->stats);
  return 0;
}

void rqb_plan_blob_free(rqb_plan_blob *b) {
  if (b && b->opaque) rqb_plan_free((rqb_plan *)b->opaque);
  if (b) memset(b, 0, sizeof(*b));
}

/* ------------------------------------------------------ HBM row matrix */
struct rqb_matrix {
  size_t rows, T, pitch;
  uint8_t *d;
  void *stream, *ev0, *ev1;
  int dev;
};
struct rqb_oplist {
  rqb_rowop *d;
  size_t n;
};

size_t rqb_matrix_pitch(const rqb_matrix *m) { return m->pitch; }

int rqb_matrix_create(rqb_matrix **out, size_t rows, size_t T) {
  *out = NULL;
  if (!rows || !T) return RQB_E_ARG;
  if (rqb_dev_count() <= 0) {
    snprintf(g_err, sizeof(g_err), "no CUDA device visible: the nanorq_b200 hot path has no CPU fallback");
    return RQB_E_NODEVICE;
  }
  rqb_matrix *m = calloc(1, sizeof(*m));
  if (!m) return RQB_E_ARG;
  m->rows = rows;
  m->T = T;
  m->pitch = round_up(T, 64);
  m->dev = rqb_dev_default();
  int e = bind_dev(m->dev);
  e = e ? e : rqb_stream_create(&m->stream);
  e = e ? e : rqb_event_create(&m->ev0);
  e = e ? e : rqb_event_create(&m->ev1);
  e = e ? e : rqb_dev_malloc((void **)&m->d, rows * m->pitch);
  e = e ? e : rqb_dev_memset(m->d, 0, rows * m->pitch, m->stream);
  if (e) {
    dev_fail(e, "rqb_matrix_create");
    rqb_matrix_destroy(m);
    return RQB_E_NODEVICE;
  }
  *out = m;
  return 0;
}

void rqb_matrix_destroy(rqb_matrix *m) {
  if (!m) return;
  if (m->stream) rqb_stream_sync(m->stream);
  if (m->d) rqb_dev_free(m->d);
  if (m->ev0) rqb_event_destroy(m->ev0);
  if (m->ev1) rqb_event_destroy(m->ev1);
  if (m->stream) rqb_stream_destroy(m->stream);
  free(m);
}

int rqb_matrix_upload(rqb_matrix *m, size_t first, size_t n, const uint8_t *src, size_t src_pitch) {
  BIND(m->dev);
  if (first + n > m->rows) return RQB_E_ARG;
  DEV(rqb_copy2d_h2d(m->d + first * m->pitch, m->pitch, src, src_pitch, m->T, n, m->stream));
  DEV(rqb_stream_sync(m->stream));
  return 0;
}

int rqb_matrix_download(rqb_matrix *m, size_t first, size_t n, uint8_t *dst, size_t dst_pitch) {
  BIND(m->dev);
  if (first + n > m->rows) return RQB_E_ARG;
  DEV(rqb_copy2d_d2h(dst, dst_pitch, m->d + first * m->pitch, m->pitch, m->T, n, m->stream));
  DEV(rqb_stream_sync(m->stream));
  return 0;
}

int rqb_matrix_fill_random(rqb_matrix *m, uint64_t seed) {
  BIND(m->dev);
  /* benchmark filler: a host xorshift tile repeated over the matrix by device copies */
  size_t tile = (size_t)1 << 22, total = m->rows * m->pitch;
  if (tile > total) tile = total;
  uint8_t *h = malloc(tile);
  uint64_t x = seed ? seed : 88172645463325252ULL;
  for (size_t k = 0; k + 8 <= tile; k += 8) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    memcpy(h + k, &x, 8);
  }
  int e = rqb_copy_h2d(m->d, h, tile, m->stream);
  e = e ? e : rqb_stream_sync(m->stream);
  free(h);
  if (e) return dev_fail(e, "fill_random");
  for (size_t done = tile; done < total; done *= 2) {
    size_t n = done < total - done ? done : total - done;
    DEV(rqb_copy_d2d(m->d + done, m->d, n, m->stream));
  }
  DEV(rqb_stream_sync(m->stream));
  return 0;
}

int rqb_ops_upload(rqb_oplist **out, const rqb_op *ops, size_t n) {
  *out = NULL;
  rqb_oplist *l = calloc(1, sizeof(*l));
  if (!l) return RQB_E_ARG;
  l->n = n;
  int e = rqb_dev_malloc((void **)&l->d, n * sizeof(rqb_rowop));
  e = e ? e : rqb_copy_h2d(l->d, ops, n * sizeof(rqb_rowop), NULL);
  e = e ? e : rqb_dev_sync();
  if (e) {
    free(l);
    return dev_fail(e, "rqb_ops_upload");
  }
  *out = l;
  return 0;
}

void rqb_ops_free(rqb_oplist *l) {
  if (!l) return;
  rqb_dev_free(l->d);
  free(l);
}

int rqb_rowops_apply_dev(rqb_matrix *m, const rqb_oplist *l, int repeats, float *ms_total) {
  BIND(m->dev);
  if (repeats < 1) repeats = 1;
  DEV(rqb_event_record(m->ev0, m->stream));
  for (int r = 0; r < repeats; r++)
    DEV(rqb_launch_rowops(m->d, m->pitch, (uint32_t)round_up(m->T, 16), l->d, (uint32_t)l->n, m->stream));
  DEV(rqb_event_record(m->ev1, m->stream));
  DEV(rqb_event_sync(m->ev1));
  if (ms_total) DEV(rqb_event_elapsed_ms(m->ev0, m->ev1, ms_total));
  return 0;
}

int rqb_rowops_apply(rqb_matrix *m, const rqb_op *ops, size_t n) {
  if (!n) return 0;
  for (size_t k = 0; k < n; k++)
    if (ops[k].i >= m->rows || (ops[k].beta && ops[k].j >= m->rows)) return RQB_E_ARG;
  rqb_oplist *l = NULL;
  int rc = rqb_ops_upload(&l, ops, n);
  if (rc) return rc;
  rc = rqb_rowops_apply_dev(m, l, 1, NULL);
  rqb_ops_free(l);
  return rc;
}

/* ------------------------------------------- reference-schedule replay */
/* the applied sequence of precode_matrix_apply_sched (lib/precode.c:23-32) and the two
 * cycle-walk permutations (lib/precode.c:3-13,381-386) composed into one gather map */
static int replay_prepare(size_t nr, const rqb_op *ops, size_t nops, long m0, long m1, const int *di,
                          size_t rows, const int *c, size_t cols, rqb_rowop **seq_out, size_t *napp_out,
                          uint32_t **map_out) {
  if (rows > nr || cols > nr || m0 < -1 || m1 < 0 || (size_t)m1 > nops || m0 >= (long)nops)
    return RQB_E_ARG;
  size_t napp = nops + 2 * (size_t)(m0 + 1), k = 0;
  for (int pass = 0; pass < 2; pass++) { /* entries of the permutations: negative = fixed point, else a row */
    const int *P = pass ? c : di;
    for (size_t i = 0; i < (pass ? cols : rows); i++)
      if (P[i] >= 0 && (size_t)P[i] >= (pass ? cols : rows)) return RQB_E_ARG;
  }
  rqb_rowop *seq = malloc(sizeof(rqb_rowop) * (napp ? napp : 1));
  if (!seq) return RQB_E_ARG;
  for (long q = 0; q < m1; q++) memcpy(&seq[k++], &ops[q], sizeof(rqb_op));
  for (long q = m0; q >= 0; q--) memcpy(&seq[k++], &ops[q], sizeof(rqb_op));
  for (long q = m1; q < (long)nops; q++) memcpy(&seq[k++], &ops[q], sizeof(rqb_op));
  for (long q = 0; q <= m0; q++) memcpy(&seq[k++], &ops[q], sizeof(rqb_op));
  for (size_t q = 0; q < napp; q++)
    if (seq[q].i >= nr || (seq[q].beta && seq[q].j >= nr)) {
      free(seq);
      return RQB_E_ARG;
    }
  uint32_t *map = malloc(4 * (nr ? nr : 1));
  if (!map) {
    free(seq);
    return RQB_E_ARG;
  }
  for (size_t r = 0; r < nr; r++) map[r] = (uint32_t)r;
  for (int pass = 0; pass < 2; pass++) {
    size_t n = pass ? cols : rows;
    int *P = malloc(sizeof(int) * (n ? n : 1));
    if (!P) {
      free(seq);
      free(map);
      return RQB_E_ARG;
    }
    memcpy(P, pass ? c : di, sizeof(int) * n);
    for (size_t i = 0; i < n; i++) {
      size_t at = i;
      while (P[at] >= 0) {
        uint32_t t = map[i];
        map[i] = map[(size_t)P[at]];
        map[(size_t)P[at]] = t;
        int nx = P[at];
        P[at] = -1;
        at = (size_t)nx;
      }
    }
    free(P);
  }
  *seq_out = seq;
  *napp_out = napp;
  *map_out = map;
  return 0;
}

/* host only: the program rqb_schedule_replay would run, for a matrix of nrows rows laid out
 * [rows | gathered rows | ZERO] (tests run it on the CPU interpreter) */
int rqb_schedule_plan_blob(size_t nrows, const rqb_op *ops, size_t nops, long m0, long m1, const int *di, size_t rows,
                           const int *c, size_t cols, rqb_plan_blob *out) {
  rqb_rowop *seq = NULL;
  uint32_t *map = NULL;
  size_t napp = 0;
  memset(out, 0, sizeof(*out));
  int rc = replay_prepare(nrows, ops, nops, m0, m1, di, rows, c, cols, &seq, &napp, &map);
  if (rc) return rc;
  rqb_plan *p = NULL;
  rc = rqb_plan_from_schedule(seq, napp, (uint32_t)nrows, map, 0, (uint32_t)nrows, (uint32_t)(2 * nrows), &p);
  free(seq);
  free(map);
  if (rc) return rc == -4 ? RQB_E_TOOBIG : RQB_E_ARG;
  out->n_pages = p->n_pages;
  out->page_bytes = RQB_PAGE_BYTES;
  out->row0[RQB_SP_IN] = 0;                     /* the matrix rows, updated in place  */
  out->row0[RQB_SP_SYM] = (uint32_t)nrows;      /* (no emitted symbols)               */
  out->row0[RQB_SP_C] = (uint32_t)nrows;        /* the gathered result                */
  out->zero_row = (uint32_t)(2 * nrows);
  out->row0[RQB_SP_WS] = out->zero_row + 1;
  out->n_rows = out->zero_row + 1;
  out->pages = p->pages;
  out->opaque = p;
  fill_stats(p, &out->stats);
  return 0;
}

/* ONE launch: the schedule becomes a program for the solve kernel (rqb_plan_from_schedule:
 * dependency levels inside the kernel, accumulations merged into gathers).  The matrix is
 * copied into a scratch arena [rows | gathered rows | ZERO], replayed there and copied back. */
int rqb_schedule_replay(rqb_matrix *m, const rqb_op *ops, size_t nops, long m0, long m1, const int *di,
                        size_t rows, const int *c, size_t cols, float *ms_device) {
  BIND(m->dev);
  rqb_rowop *seq = NULL;
  uint32_t *map = NULL;
  size_t napp = 0;
  int rc = replay_prepare(m->rows, ops, nops, m0, m1, di, rows, c, cols, &seq, &napp, &map);
  if (rc) return rc;
  const size_t nr = m->rows;
  rqb_plan *plan = NULL;
  rc = rqb_plan_from_schedule(seq, napp, (uint32_t)nr, map, 0, (uint32_t)nr, (uint32_t)(2 * nr), &plan);
  free(seq);
  free(map);
  if (rc) {
    snprintf(g_err, sizeof(g_err), "rqb_plan_from_schedule failed (%d)", rc);
    return rc == -4 ? RQB_E_TOOBIG : RQB_E_ARG;
  }
  uint8_t *arena = NULL, *d_pages = NULL;
  rqb_solve_args *d_args = NULL, h_args;
  const size_t pb = (size_t)plan->n_pages * RQB_PAGE_BYTES;
  int e = rqb_dev_malloc((void **)&arena, (2 * nr + 1) * m->pitch);
  e = e ? e : rqb_dev_malloc((void **)&d_pages, pb);
  e = e ? e : rqb_dev_malloc((void **)&d_args, sizeof(h_args));
  memset(&h_args, 0, sizeof(h_args));
  h_args.base = arena;
  h_args.pages = d_pages;
  h_args.pitch = (uint32_t)m->pitch;
  h_args.n_pages = plan->n_pages;
  h_args.width = (uint32_t)round_up(m->T, 16);
  e = e ? e : rqb_copy_h2d(d_pages, plan->pages, pb, m->stream); /* pageable source: staged before the call returns */
  e = e ? e : rqb_copy_h2d(d_args, &h_args, sizeof(h_args), m->stream);
  e = e ? e : rqb_dev_memset(arena + 2 * nr * m->pitch, 0, m->pitch, m->stream);
  e = e ? e : rqb_event_record(m->ev0, m->stream);
  e = e ? e : rqb_copy_d2d(arena, m->d, nr * m->pitch, m->stream);
  e = e ? e : rqb_launch_solve(d_args, 1, h_args.width, m->stream);
  e = e ? e : rqb_copy_d2d(m->d, arena + nr * m->pitch, nr * m->pitch, m->stream);
  e = e ? e : rqb_event_record(m->ev1, m->stream);
  e = e ? e : rqb_stream_sync(m->stream);
  if (!e && ms_device) e = rqb_event_elapsed_ms(m->ev0, m->ev1, ms_device);
  if (arena) rqb_dev_free(arena);
  if (d_pages) rqb_dev_free(d_pages);
  if (d_args) rqb_dev_free(d_args);
  rqb_plan_free(plan);
  if (e) return dev_fail(e, "rqb_schedule_replay");
  return 0;
}

/* The literal form: every dependency level of the op list is one launch of the batched
 * row-op kernel (oaxpy/oaddrow/oscal as they are), then the out-of-place gather. */
int rqb_schedule_replay_stepwise(rqb_matrix *m, const rqb_op *ops, size_t nops, long m0, long m1, const int *di,
                                 size_t rows, const int *c, size_t cols, float *ms_device) {
  BIND(m->dev);
  rqb_rowop *seq = NULL;
  uint32_t *map = NULL;
  size_t napp = 0;
  int rc = replay_prepare(m->rows, ops, nops, m0, m1, di, rows, c, cols, &seq, &napp, &map);
  if (rc) return rc;
  /* levelise: an op runs after every earlier op that wrote one of its rows or
   * read its destination; ops of one level are then mutually independent */
  uint32_t *lw = calloc(m->rows, 4), *lr = calloc(m->rows, 4), *lev = malloc(4 * (napp ? napp : 1));
  uint32_t *start = NULL, *cur = NULL;
  rqb_rowop *sorted = malloc(sizeof(rqb_rowop) * (napp ? napp : 1));
  uint32_t nlev = 0;
  if (!lw || !lr || !lev || !sorted) {
  nomem:
    free(seq); free(lw); free(lr); free(lev); free(start); free(cur); free(sorted); free(map);
    return RQB_E_ARG;
  }
  for (size_t q = 0; q < napp; q++) {
    uint32_t i = seq[q].i, l = lw[i] > lr[i] ? lw[i] : lr[i];
    if (seq[q].beta) {
      uint32_t j = seq[q].j;
      if (lw[j] > l) l = lw[j];
      l++;
      if (lr[j] < l) lr[j] = l;
    } else {
      l++;
    }
    lw[i] = l;
    lev[q] = l;
    if (l > nlev) nlev = l;
  }
  start = calloc((size_t)nlev + 2, 4);
  cur = malloc(4 * ((size_t)nlev + 2));
  if (!start || !cur) goto nomem;
  for (size_t q = 0; q < napp; q++) start[lev[q] + 1]++;
  for (uint32_t l = 0; l <= nlev; l++) start[l + 1] += start[l];
  memcpy(cur, start, 4 * ((size_t)nlev + 2));
  for (size_t q = 0; q < napp; q++) sorted[cur[lev[q]]++] = seq[q];
  free(cur);
  cur = NULL;
  size_t nr = m->rows;
  rqb_rowop *d_ops = NULL;
  uint32_t *d_map = NULL;
  uint8_t *d_tmp = NULL;
  int e = rqb_dev_malloc((void **)&d_ops, sizeof(rqb_rowop) * (napp ? napp : 1));
  e = e ? e : rqb_dev_malloc((void **)&d_map, 4 * nr);
  e = e ? e : rqb_dev_malloc((void **)&d_tmp, nr * m->pitch);
  e = e ? e : rqb_copy_h2d(d_ops, sorted, sizeof(rqb_rowop) * napp, m->stream);
  e = e ? e : rqb_copy_h2d(d_map, map, 4 * nr, m->stream);
  e = e ? e : rqb_event_record(m->ev0, m->stream);
  uint32_t width = (uint32_t)round_up(m->T, 16);
  for (uint32_t l = 1; l <= nlev && !e; l++)
    e = rqb_launch_rowops(m->d, m->pitch, width, d_ops + start[l], start[l + 1] - start[l], m->stream);
  e = e ? e : rqb_launch_gather_rows(d_tmp, m->pitch, m->d, m->pitch, d_map, (uint32_t)nr, width, m->stream);
  e = e ? e : rqb_copy_d2d(m->d, d_tmp, nr * m->pitch, m->stream);
  e = e ? e : rqb_event_record(m->ev1, m->stream);
  e = e ? e : rqb_stream_sync(m->stream);
  if (!e && ms_device) e = rqb_event_elapsed_ms(m->ev0, m->ev1, ms_device);
  if (d_ops) rqb_dev_free(d_ops);
  if (d_map) rqb_dev_free(d_map);
  if (d_tmp) rqb_dev_free(d_tmp);
  free(seq); free(lw); free(lr); free(lev); free(start); free(sorted); free(map);
  if (e) return dev_fail(e, "rqb_schedule_replay_stepwise");
  return 0;
}
