// h2d_ring_test.cu -- is a small cache-resident pinned ring a cheaper way to the GPU than a big
// pinned staging area?   nvcc -O2 -Xcompiler -mavx2,-pthread tools/h2d_ring_test.cu -o /tmp/h2d_ring && /tmp/h2d_ring 16
// Every thread moves `blocks` blocks of 4096 rows of 1280 bytes from pageable memory to its own
// device buffer, (a) through a block-sized pinned staging area written with streaming stores and
// ONE cudaMemcpyAsync (what nanorq_api.c does today), (b) through a ring of 8 pinned chunks of
// 128 rows written with ordinary stores, one cudaMemcpyAsync per chunk.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <pthread.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static const size_t T = 1280, K = 4096, CH = 128, NCH = 8;
static int MODE, BLOCKS = 24;
static pthread_barrier_t bar;
static void nt(char *d, const char *s, size_t n) {
  for (size_t k = 0; k < n; k += 64) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(s + k)), b = _mm256_loadu_si256((const __m256i *)(s + k + 32));
    _mm256_stream_si256((__m256i *)(d + k), a); _mm256_stream_si256((__m256i *)(d + k + 32), b);
  }
}
static void *work(void *) {
  char *src = (char *)malloc(K * T), *pin, *dev;
  memset(src, 1, K * T);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaMalloc((void **)&dev, K * T);
  cudaMallocHost((void **)&pin, MODE ? CH * NCH * T : K * T);
  memset(pin, 0, MODE ? CH * NCH * T : K * T);
  pthread_barrier_wait(&bar);
  for (int b = 0; b < BLOCKS; b++) {
    if (MODE == 0) {
      for (size_t r = 0; r < K; r++) nt(pin + r * T, src + r * T, T);
      _mm_sfence();
      cudaMemcpyAsync(dev, pin, K * T, cudaMemcpyHostToDevice, st);
      cudaStreamSynchronize(st);
    } else {
      for (size_t c = 0; c < K / CH; c++) {
        char *slot = pin + (c % NCH) * CH * T;
        if (c && c % NCH == 0) cudaStreamSynchronize(st); // the ring wraps: everything queued has been read
        for (size_t r = 0; r < CH; r++) memcpy(slot + r * T, src + (c * CH + r) * T, T);
        cudaMemcpyAsync(dev + c * CH * T, slot, CH * T, cudaMemcpyHostToDevice, st);
      }
      cudaStreamSynchronize(st);
    }
  }
  pthread_barrier_wait(&bar);
  return nullptr;
}
int main(int argc, char **argv) {
  int n = argc > 1 ? atoi(argv[1]) : 16;
  cudaFree(0);
  for (int rep = 0; rep < 2; rep++)
    for (MODE = 0; MODE < 2; MODE++) {
      pthread_t th[128];
      pthread_barrier_init(&bar, nullptr, n + 1);
      for (int k = 0; k < n; k++) pthread_create(&th[k], nullptr, work, nullptr);
      pthread_barrier_wait(&bar); double t0 = now_s(); pthread_barrier_wait(&bar); double t1 = now_s();
      for (int k = 0; k < n; k++) pthread_join(th[k], nullptr);
      printf("%s, %d threads: %.1f GB/s of payload to the device (%.2f ms per 5.2 MB block per thread)\n",
             MODE ? "ring of 8 x 128-row pinned chunks, ordinary stores" : "block-sized pinned staging, streaming stores    ", n,
             n * (double)BLOCKS * K * T / (t1 - t0) / 1e9, 1e3 * (t1 - t0) / BLOCKS);
    }
  return 0;
}
