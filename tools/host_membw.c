/* host_membw.c -- how fast can N threads copy 1280-byte rows on this host, with
 * ordinary stores (glibc memcpy) and with non-temporal stores?  Sizing aid for
 * the nanorq.h e2e path, which moves tens of MB per K=4096 block through host
 * memory.
 *   gcc -O2 -mavx2 -pthread tools/host_membw.c -o /tmp/host_membw && /tmp/host_membw 1 4 8 16 */
#define _POSIX_C_SOURCE 200809L
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static size_t BYTES = 64u << 20;
static int REPS = 8, MODE = 0;
static pthread_barrier_t bar;
static void copy_nt(char *d, const char *s, size_t n) {
  for (size_t k = 0; k + 32 <= n; k += 32)
    _mm256_stream_si256((__m256i *)(d + k), _mm256_loadu_si256((const __m256i *)(s + k)));
  _mm_sfence();
}
static void *work(void *a) {
  (void)a;
  char *src = aligned_alloc(64, BYTES), *dst = aligned_alloc(64, BYTES);
  memset(src, 1, BYTES); memset(dst, 2, BYTES);
  pthread_barrier_wait(&bar);
  for (int r = 0; r < REPS; r++)
    for (size_t o = 0; o + 1280 <= BYTES; o += 1280) {
      if (MODE) copy_nt(dst + o, src + o, 1280); else memcpy(dst + o, src + o, 1280);
    }
  pthread_barrier_wait(&bar);
  free(src); free(dst);
  return NULL;
}
int main(int argc, char **argv) {
  for (MODE = 0; MODE < 2; MODE++)
  for (int a = 1; a < argc; a++) {
    int n = atoi(argv[a]);
    pthread_t th[256];
    pthread_barrier_init(&bar, NULL, n + 1);
    for (int k = 0; k < n; k++) pthread_create(&th[k], NULL, work, NULL);
    pthread_barrier_wait(&bar);
    double t0 = now_s();
    pthread_barrier_wait(&bar);
    double t1 = now_s();
    for (int k = 0; k < n; k++) pthread_join(th[k], NULL);
    printf("%s threads %3d: %.1f GB/s copied (%.1f GB/s per thread), 1280-byte rows\n", MODE ? "nt-store" : "memcpy  ", n,
           n * (double)BYTES * REPS / (t1 - t0) / 1e9, (double)BYTES * REPS / (t1 - t0) / 1e9);
    pthread_barrier_destroy(&bar);
  }
  return 0;
}
