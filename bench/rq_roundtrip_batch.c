/* rq_roundtrip_batch.c -- the round trips of rq_roundtrip.c through the BATCH entry
 * points of this library (nanorq_batch.h): nanorq_encode_range and
 * nanorq_decoder_add_symbols over page-locked buffers, ioctx_from_pinned_mem on both
 * ends.  Same workload, same seeds, same loss patterns and the same verification as
 * rq_roundtrip.c (which stays the apples-to-apples arm: one source, linked against
 * both this library and the unmodified reference); this file only exists for this
 * library because the reference has no batch calls.
 *
 * Per block the host does: two object constructors, one encode-range call (the block's
 * source symbols and the repair symbols the receiver will need, into a packet ring whose
 * row number is the sequence number), one add-symbols call over that ring with the lost
 * packets marked as holes (NANORQ_TAG_NONE), one repair call.  No symbol byte is copied by the CPU:
 * payload -> device, device -> packet buffer, packet buffer -> device, device -> output
 * are all DMA.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nanorq_batch.h"
#include "rqb200.h"

typedef struct {
  int K, T, nblocks;
  double loss;
  int overhead;
  unsigned seed;
  int nthreads, precalc, verify;
} rt_config; /* same layout as rq_roundtrip.c */

typedef struct {
  double wall_s, t_gen, t_emit, t_add, t_repair;
  long n_lost, n_sent;
  int retries, failures, mismatches;
  unsigned long long out_fnv;
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
  uint8_t **payload, **decoded; /* page-locked */
  uint8_t *pk;                  /* this worker's page-locked packet ring */
  int from_payload, no_decode;
  int *next;
  pthread_mutex_t *mu;
  pthread_barrier_t *bar;
  rt_result acc;
  double t_start, t_done;
} worker;

static int take(worker *w) {
  pthread_mutex_lock(w->mu);
  int b = *w->next < w->cfg->nblocks ? (*w->next)++ : -1;
  pthread_mutex_unlock(w->mu);
  return b;
}

static void *work(void *arg) {
  worker *w = arg;
  const rt_config *c = w->cfg;
  const size_t K = (size_t)c->K, T = (size_t)c->T, F = K * T;
  const size_t max_pk = 2 * K + (size_t)c->overhead + 64;
  uint8_t *pk = w->pk; /* the page-locked packet ring a sender/receiver would reuse */
  uint32_t *tags = malloc(max_pk * sizeof(uint32_t));
  pthread_barrier_wait(w->bar);
  w->t_start = now_s();
  for (int b; pk && (b = take(w)) >= 0;) {
    uint32_t rs = (c->seed + 0x9e3779b9u * (uint32_t)(b + 1)) | 1u;
    const uint32_t thresh = (uint32_t)(c->loss * 4294967296.0 > 4294967295.0 ? 4294967295.0 : c->loss * 4294967296.0);
    nanorq *enc = nanorq_encoder_new_ex(F, (uint16_t)T, (uint16_t)K, 0, 8);
    if (!enc) { w->acc.failures++; continue; }
    struct ioctx *in = ioctx_from_pinned_mem(w->payload[b], F, 1);
    if (c->precalc) nanorq_precalculate(enc);
    double t0 = now_s();
    bool ok = nanorq_generate_symbols(enc, 0, in);
    double t1 = now_s();
    /* the sender emits the whole block and the repair symbols with ONE call into its packet ring
     * (row = sequence number); the channel drops packets: their rows become holes */
    size_t lost = 0;
    for (uint32_t esi = 0; esi < K; esi++) {
      const int drop = xs32(&rs) < thresh;
      lost += (size_t)drop;
      tags[esi] = drop ? NANORQ_TAG_NONE : nanorq_tag(0, esi);
    }
    const uint32_t n_rep = (uint32_t)(lost + (size_t)c->overhead);
    size_t n = K + n_rep;
    uint32_t next_rep = (uint32_t)(K + n_rep);
    for (uint32_t q = 0; q < n_rep; q++) tags[K + q] = nanorq_tag(0, (uint32_t)K + q);
    /* RQ_BATCH_SENDER=payload: the sender transmits source symbols straight from its payload buffer
     * (they are the payload) and asks the encoder for the repair symbols only */
    const int from_payload = w->from_payload;
    if (ok && !from_payload) ok = nanorq_encode_range(enc, 0, 0, (uint32_t)n, pk, T, in) == n;
    if (ok && from_payload && n_rep) ok = nanorq_encode_range(enc, 0, (uint32_t)K, n_rep, pk + K * T, T, in) == n_rep;
    double t2 = now_s();
    uint64_t oti_c = nanorq_oti_common(enc);
    uint32_t oti_s = nanorq_oti_scheme_specific(enc);
    nanorq *dec = ok ? nanorq_decoder_new(oti_c, oti_s) : NULL;
    struct ioctx *out = ioctx_from_pinned_mem(w->decoded[b], F, 1);
    bool done = false;
    double t3 = t2, t4 = t2;
    if (dec) {
      if (w->no_decode) {
        nanorq_free(dec);
        dec = NULL;
        memcpy(w->decoded[b], w->payload[b], F); /* experiment: encoder side only */
        w->acc.t_gen += t1 - t0;
        w->acc.t_emit += t2 - t1;
        out->destroy(out);
        in->destroy(in);
        nanorq_free(enc);
        continue;
      }
      if (!from_payload) {
        if (nanorq_decoder_add_symbols(dec, tags, pk, T, n, NULL, out) < 0) ok = false;
      } else { /* source symbols from the payload buffer (with holes), repair symbols from the ring */
        if (nanorq_decoder_add_symbols(dec, tags, w->payload[b], T, K, NULL, out) < 0) ok = false;
        if (n_rep && nanorq_decoder_add_symbols(dec, tags + K, pk + K * T, T, n_rep, NULL, out) < 0) ok = false;
      }
      t3 = now_s();
      done = ok && nanorq_repair_block(dec, out, 0);
      t4 = now_s();
      for (int retry = 0; ok && !done && retry < 8; retry++) {
        if (retry == 0) w->acc.retries++;
        if (n + 2 > max_pk) break;
        tags[n] = nanorq_tag(0, next_rep);
        tags[n + 1] = nanorq_tag(0, next_rep + 1);
        if (nanorq_encode_range(enc, 0, next_rep, 2, pk + n * T, T, in) != 2) ok = false;
        nanorq_decoder_add_symbols(dec, tags + n, pk + n * T, T, 2, NULL, out);
        n += 2;
        next_rep += 2;
        done = ok && nanorq_repair_block(dec, out, 0);
        t4 = now_s();
      }
    }
    if (!done) w->acc.failures++;
    w->acc.t_gen += t1 - t0;
    w->acc.t_emit += t2 - t1;
    w->acc.t_add += t3 - t2;
    w->acc.t_repair += t4 - t3;
    w->acc.n_lost += (long)lost;
    w->acc.n_sent += (long)n;
    if (dec) nanorq_free(dec);
    out->destroy(out);
    in->destroy(in);
    nanorq_free(enc);
  }
  w->t_done = now_s();
  if (!pk) w->acc.failures++;
  free(tags);
  return NULL;
}

int rq_roundtrip_batch_run(const rt_config *cfg, rt_result *res) {
  memset(res, 0, sizeof(*res));
  if (cfg->K < 1 || cfg->T < 1 || cfg->nblocks < 1 || cfg->nthreads < 1) return -1;
  const size_t F = (size_t)cfg->K * (size_t)cfg->T;
  const int nb = cfg->nblocks, nt = cfg->nthreads < nb ? cfg->nthreads : nb;
  /* payloads and outputs live in two page-locked arenas (allocated and faulted in before the
   * clock starts, like the malloc'ed buffers of rq_roundtrip.c) */
  uint8_t *pay = rqb_host_alloc((size_t)nb * (F + 64)), *dec = rqb_host_alloc((size_t)nb * (F + 64));
  uint8_t **payload = calloc((size_t)nb, sizeof(*payload)), **decoded = calloc((size_t)nb, sizeof(*decoded));
  if (!pay || !dec || !payload || !decoded) return -2;
  for (int b = 0; b < nb; b++) {
    payload[b] = pay + (size_t)b * (F + 64);
    decoded[b] = dec + (size_t)b * (F + 64);
    memset(decoded[b], 0, F);
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
  /* page-locked memory is allocated and released outside the clock and while no worker runs:
   * cudaMallocHost / cudaFreeHost stall every thread of the process that talks to the driver */
  const size_t pk_bytes = (2 * (size_t)cfg->K + (size_t)cfg->overhead + 64) * (size_t)cfg->T;
  for (int k = 0; k < nt; k++) {
    ws[k].pk = rqb_host_alloc(pk_bytes);
    if (ws[k].pk) memset(ws[k].pk, 0, pk_bytes);
  }
  for (int k = 0; k < nt; k++) {
    ws[k].cfg = cfg;
    ws[k].from_payload = getenv("RQ_BATCH_SENDER") && !strcmp(getenv("RQ_BATCH_SENDER"), "payload");
    ws[k].no_decode = getenv("RQ_BATCH_NO_DECODE") != NULL;
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
    if (ws[k].t_start < t0) t0 = ws[k].t_start;
    if (ws[k].t_done > t_end) t_end = ws[k].t_done;
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
  res->out_fnv = h;
  pthread_barrier_destroy(&bar);
  for (int k = 0; k < nt; k++) rqb_host_release(ws[k].pk);
  rqb_host_release(pay);
  rqb_host_release(dec);
  free(payload);
  free(decoded);
  free(ws);
  free(th);
  return 0;
}
