/* pinned_read.c -- is reading cudaMallocHost memory with the CPU as fast as reading malloc memory?
 *   nvcc -O2 tools/pinned_read.c -o /tmp/pinned_read -Xcompiler -mavx2 && /tmp/pinned_read */
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void nt(char *d, const char *s, size_t n) {
  for (size_t k = 0; k < n; k += 64) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(s + k)), b = _mm256_loadu_si256((const __m256i *)(s + k + 32));
    _mm256_stream_si256((__m256i *)(d + k), a); _mm256_stream_si256((__m256i *)(d + k + 32), b);
  }
}
static double run(char *dst, const char *src, size_t B, int mode) {
  double best = 1e9;
  for (int rep = 0; rep < 12; rep++) {
    double t0 = now_s();
    for (size_t o = 0; o + 1280 <= B; o += 1280) { if (mode) nt(dst + o, src + o, 1280); else memcpy(dst + o, src + o, 1280); }
    _mm_sfence();
    double t = now_s() - t0;
    if (t < best) best = t;
  }
  return B / best / 1e9;
}
int main(void) {
  size_t B = 4096 * 1280 * 4;
  char *pin, *pin2, *m1 = (char *)aligned_alloc(64, B), *m2 = (char *)aligned_alloc(64, B);
  cudaMallocHost((void **)&pin, B); cudaMallocHost((void **)&pin2, B);
  memset(pin, 1, B); memset(pin2, 1, B); memset(m1, 2, B); memset(m2, 3, B);
  printf("memcpy  malloc->malloc %.2f GB/s | pinned->malloc %.2f | malloc->pinned %.2f | pinned->pinned %.2f\n", run(m2, m1, B, 0), run(m2, pin, B, 0), run(pin, m1, B, 0), run(pin2, pin, B, 0));
  printf("nt      malloc->malloc %.2f GB/s | pinned->malloc %.2f | malloc->pinned %.2f | pinned->pinned %.2f\n", run(m2, m1, B, 1), run(m2, pin, B, 1), run(pin, m1, B, 1), run(pin2, pin, B, 1));
  return 0;
}
