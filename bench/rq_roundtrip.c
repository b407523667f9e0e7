/* rq_roundtrip.c -- encode -> erase -> decode round trips through the PUBLIC
 * nanorq API (nanorq.h + io.h) only.
 *
 * The same source is compiled twice: against this repository's
 * libnanorq_b200.so (nanorq_b200/librq_roundtrip.so, built by
 * nanorq_b200/build.py) and against the unmodified reference
 * (oracle/_ref/librq_roundtrip_ref.so, built by oracle/Makefile).  That the one
 * file links against both is the drop-in check; bench.py times both builds on
 * the same seeded workload (its e2e figure and its --impl reference arm).
 *
 * Workload, after the reference's benchmark.c:82-170 but reproducible: every
 * block is an object of F = K*T bytes (nanorq_encoder_new_ex(F,T,K,0,8)) filled
 * by xorshift32(seed+block); each source symbol is dropped with probability
 * `loss` (xorshift PRNG seeded per block); the receiver gets the surviving
 * source symbols in ESI order followed by repair symbols K, K+1, ... (as many
 * as were dropped, plus `overhead`).  If the decoder reports "need more"
 * (singular matrix) two further repair symbols are fed and the repair retried.
 *
 * Objects: the blocks are grouped into objects of `zblocks` source blocks each
 * (nanorq_encoder_new_ex(Z*K*T, T, K, 0, 8) => Z blocks of K symbols), the way an
 * application sends a large object; nanorq_precalculate is called once per object so
 * that an implementation which keeps a per-object schedule (the reference: rq->S,
 * lib/nanorq.c:219-221,393-401) amortises it over the object's blocks.  zblocks = 1
 * gives one object per block.  Per-block state is released with
 * nanorq_encoder_cleanup as soon as a block is done.
 * Threads: `nthreads` workers take objects from a shared counter; no nanorq object is
 * shared between threads.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nanorq.h"

typedef struct {
  int K;
  int T;
  int nblocks;
  double loss;
  int overhead;
  unsigned seed;
  int nthreads;
  int precalc; /* call nanorq_precalculate before generate_symbols */
  int verify;  /* compare decoded bytes with the payload (after the clock stops) */
  int zblocks; /* source blocks per object (0 or 1: one object per block) */
} rt_config;

typedef struct {
  double wall_s;   /* first worker out of the start barrier -> last worker done (payload generation and verification excluded) */
  double t_gen;    /* sums over blocks of the time inside ...   nanorq_generate_symbols */
  double t_emit;   /*                                           nanorq_encode (all symbols sent) */
  double t_add;    /*                                           nanorq_decoder_new + add_symbol  */
  double t_repair; /*                                           nanorq_repair_block              */
  long n_lost;     /* source symbols dropped over all blocks */
  long n_sent;     /* symbols fed to decoders */
  int retries;     /* blocks that needed extra repair symbols */
  int failures;    /* blocks that never decoded */
  int mismatches;  /* blocks whose decoded bytes differ from the payload */
  unsigned long long out_fnv; /* FNV-1a-64 over all decoded bytes, block order */
} rt_result;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline uint32_t xs32(uint32_t *s) {
  uint32_t x = *s;
  x ^= x << 13;
  x ^= x >> 17;
  x ^= x << 5;
  return *s = x;
}

typedef struct {
  const rt_config *cfg;
  uint8_t **payload, **decoded;
  int *next;
  pthread_mutex_t *mu;
  pthread_barrier_t *bar;
  rt_result acc;
  double t_start, t_done;
} worker;

static int take(worker *w, int nobj) {
  pthread_mutex_lock(w->mu);
  int b = *w->next < nobj ? (*w->next)++ : -1;
  pthread_mutex_unlock(w->mu);
  return b;
}

static void *work(void *arg) {
  worker *w = arg;
  const rt_config *c = w->cfg;
  const size_t K = (size_t)c->K, T = (size_t)c->T, F = K * T;
  const int Z = c->zblocks > 1 ? c->zblocks : 1, nobj = (c->nblocks + Z - 1) / Z;
  const size_t max_pk = 2 * K + (size_t)c->overhead + 64;
  uint8_t *pk = malloc(max_pk * T);
  uint32_t *tags = malloc(max_pk * sizeof(uint32_t));
  memset(pk, 0, max_pk * T); /* fault the packet buffer in before the clock starts */
  pthread_barrier_wait(w->bar);
  w->t_start = now_s();
  for (int o; (o = take(w, nobj)) >= 0;) {
    const int b0 = o * Z, zb = c->nblocks - b0 < Z ? c->nblocks - b0 : Z; /* blocks b0 .. b0+zb-1 are this object */
    const uint32_t thresh = (uint32_t)(c->loss * 4294967296.0 > 4294967295.0 ? 4294967295.0 : c->loss * 4294967296.0);
    /* ---- sender and receiver objects (payload and decoded output of an object are contiguous) */
    nanorq *enc = nanorq_encoder_new_ex(F * (size_t)zb, (uint16_t)T, (uint16_t)K, 0, 8);
    if (!enc || nanorq_blocks(enc) != (size_t)zb) {
      w->acc.failures += zb;
      if (enc) nanorq_free(enc);
      continue;
    }
    struct ioctx *in = ioctx_from_mem(w->payload[b0], F * (size_t)zb);
    if (c->precalc) nanorq_precalculate(enc);
    nanorq *dec = nanorq_decoder_new(nanorq_oti_common(enc), nanorq_oti_scheme_specific(enc));
    struct ioctx *out = ioctx_from_mem(w->decoded[b0], F * (size_t)zb);
    if (!dec) {
      w->acc.failures += zb;
      in->destroy(in);
      out->destroy(out);
      nanorq_free(enc);
      continue;
    }
    for (int z = 0; z < zb; z++) {
      const uint8_t sbn = (uint8_t)z;
      uint32_t rs = (c->seed + 0x9e3779b9u * (uint32_t)(b0 + z + 1)) | 1u;
      double t0 = now_s();
      bool ok = nanorq_generate_symbols(enc, sbn, in);
      double t1 = now_s();
      size_t n = 0, lost = 0;
      for (uint32_t esi = 0; esi < K && ok; esi++) {
        if (xs32(&rs) < thresh) { lost++; continue; }
        tags[n] = nanorq_tag(sbn, esi);
        ok = nanorq_encode(enc, pk + n * T, esi, sbn, in) == T;
        n++;
      }
      uint32_t next_rep = (uint32_t)K;
      for (size_t k = 0; k < lost + (size_t)c->overhead && ok; k++, n++, next_rep++) {
        tags[n] = nanorq_tag(sbn, next_rep);
        ok = nanorq_encode(enc, pk + n * T, next_rep, sbn, in) == T;
      }
      double t2 = now_s();
      /* ---- receiver */
      for (size_t k = 0; k < n; k++)
        if (nanorq_decoder_add_symbol(dec, pk + k * T, tags[k], out) == NANORQ_SYM_ERR) ok = false;
      double t3 = now_s();
      bool done = ok && nanorq_repair_block(dec, out, sbn);
      double t4 = now_s();
      for (int retry = 0; ok && !done && retry < 8; retry++) {
        if (retry == 0) w->acc.retries++;
        for (int x = 0; x < 2 && n < max_pk; x++, n++, next_rep++) {
          tags[n] = nanorq_tag(sbn, next_rep);
          if (nanorq_encode(enc, pk + n * T, next_rep, sbn, in) != T) ok = false;
          nanorq_decoder_add_symbol(dec, pk + n * T, tags[n], out);
        }
        done = ok && nanorq_repair_block(dec, out, sbn);
        t4 = now_s();
      }
      if (!done) w->acc.failures++;
      w->acc.t_gen += t1 - t0;
      w->acc.t_emit += t2 - t1;
      w->acc.t_add += t3 - t2;
      w->acc.t_repair += t4 - t3;
      w->acc.n_lost += (long)lost;
      w->acc.n_sent += (long)n;
      /* this block is finished on both sides: release its state (the encoder's per-block matrices
       * in the reference, lib/nanorq.c:437-451; a no-op for a decoder object there) */
      nanorq_encoder_cleanup(enc, sbn);
      nanorq_encoder_cleanup(dec, sbn);
    }
    nanorq_free(dec);
    out->destroy(out);
    in->destroy(in);
    nanorq_free(enc);
  }
  w->t_done = now_s();
  free(pk);
  free(tags);
  return NULL;
}

int rq_roundtrip_run(const rt_config *cfg, rt_result *res) {
  memset(res, 0, sizeof(*res));
  if (cfg->K < 1 || cfg->T < 1 || cfg->nblocks < 1 || cfg->nthreads < 1) return -1;
  const size_t F = (size_t)cfg->K * (size_t)cfg->T;
  const int nb = cfg->nblocks, nobjs = (nb + (cfg->zblocks > 1 ? cfg->zblocks : 1) - 1) / (cfg->zblocks > 1 ? cfg->zblocks : 1);
  const int nt = cfg->nthreads < nobjs ? cfg->nthreads : nobjs;
  uint8_t **payload = calloc((size_t)nb, sizeof(*payload)), **decoded = calloc((size_t)nb, sizeof(*decoded));
  /* the blocks of an object are contiguous: one arena each for payloads and decoded output */
  uint8_t *pay_arena = malloc((size_t)nb * F + 4), *dec_arena = malloc((size_t)nb * F + 4);
  for (int b = 0; b < nb; b++) {
    payload[b] = pay_arena + (size_t)b * F;
    decoded[b] = dec_arena + (size_t)b * F;
    memset(decoded[b], 0, F); /* faulted in before the clock starts, like the payload and packet buffers */
    uint32_t s = cfg->seed + 42u + (uint32_t)b;
    if (!s) s = 1;
    for (size_t k = 0; k < F; k += 4) {
      uint32_t v = xs32(&s);
      memcpy(payload[b] + k, &v, 4);
    }
  }
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)nt + 1);
  int next = 0;
  worker *ws = calloc((size_t)nt, sizeof(*ws));
  pthread_t *th = calloc((size_t)nt, sizeof(*th));
  for (int k = 0; k < nt; k++) {
    ws[k].cfg = cfg;
    ws[k].payload = payload;
    ws[k].decoded = decoded;
    ws[k].next = &next;
    ws[k].mu = &mu;
    ws[k].bar = &bar;
    pthread_create(&th[k], NULL, work, &ws[k]);
  }
  pthread_barrier_wait(&bar);
  double t0 = now_s(), t_end = 0.0;
  for (int k = 0; k < nt; k++) {
    pthread_join(th[k], NULL);
    if (ws[k].t_start < t0) t0 = ws[k].t_start; /* first worker out of the barrier ... */
    if (ws[k].t_done > t_end) t_end = ws[k].t_done; /* ... to the last one done */
    res->t_gen += ws[k].acc.t_gen;
    res->t_emit += ws[k].acc.t_emit;
    res->t_add += ws[k].acc.t_add;
    res->t_repair += ws[k].acc.t_repair;
    res->n_lost += ws[k].acc.n_lost;
    res->n_sent += ws[k].acc.n_sent;
    res->retries += ws[k].acc.retries;
    res->failures += ws[k].acc.failures;
  }
  res->wall_s = t_end - t0;
  unsigned long long h = 14695981039346656037ULL;
  for (int b = 0; b < nb; b++) {
    if (cfg->verify && memcmp(payload[b], decoded[b], F) != 0) res->mismatches++;
    for (size_t k = 0; k < F; k++) h = (h ^ decoded[b][k]) * 1099511628211ULL;
  }
  free(pay_arena);
  free(dec_arena);
  res->out_fnv = h;
  pthread_barrier_destroy(&bar);
  free(payload);
  free(decoded);
  free(ws);
  free(th);
  return 0;
}
