/* rqb_hostcopy.h -- row copies on the host side of the nanorq.h layer.
 *
 * A K=4096 block moves ~20 MB through memcpy between the caller's buffers, the
 * pinned staging rows and the output ioctx; with all host cores doing that the
 * box is memory-bandwidth bound.  Rows whose destination is not read again by
 * the CPU soon (staging rows the DMA engine picks up, decoded output) are
 * written with non-temporal stores: no read-for-ownership of the destination
 * line, one third less DRAM traffic per copy.  See rqb_copy_fence(). */
#ifndef RQB_HOSTCOPY_H
#define RQB_HOSTCOPY_H

#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline void rqb_copy_stream(void *dst, const void *src, size_t n) {
  uint8_t *d = (uint8_t *)dst;
  const uint8_t *s = (const uint8_t *)src;
  if (n < 256) {
    memcpy(d, s, n);
    return;
  }
  size_t head = (64 - ((uintptr_t)d & 63)) & 63; /* whole 64-byte lines combine in the write buffers */
  if (head) {
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
  }
  size_t body = n & ~(size_t)63;
  for (size_t k = 0; k < body; k += 64) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(s + k));
    __m256i b = _mm256_loadu_si256((const __m256i *)(s + k + 32));
    _mm256_stream_si256((__m256i *)(d + k), a);
    _mm256_stream_si256((__m256i *)(d + k + 32), b);
  }
  if (n > body) memcpy(d + body, s + body, n - body);
}

/* Non-temporal stores are weakly ordered.  A fence per 1280-byte row costs more than
 * the copy itself (measured: 2.9 ms instead of 0.9 ms per 4096 rows), so callers fence
 * ONCE where the rows are handed to someone else: before a DMA upload is queued, when
 * a block's output is complete, when an ioctx or codec object is destroyed. */
static inline void rqb_copy_fence(void) { _mm_sfence(); }

#endif
