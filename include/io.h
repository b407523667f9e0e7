/* io.h -- I/O context used by the nanorq API (same surface as the reference's
 * include/io.h:7-20 so programs written against it link unchanged). */
#ifndef NANORQ_IOCTX_H
#define NANORQ_IOCTX_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct ioctx {
  size_t (*read)(struct ioctx *, uint8_t *, size_t);
  size_t (*write)(struct ioctx *, const uint8_t *, size_t);
  bool (*seek)(struct ioctx *, const size_t);
  size_t (*size)(struct ioctx *);
  long (*tell)(struct ioctx *);
  void (*destroy)(struct ioctx *);
  bool seekable;
  bool writable;
};

/* t = 1: open for reading (encoder side), t = 0: create (decoder side) */
struct ioctx *ioctx_from_file(const char *fn, int t);
struct ioctx *ioctx_mmap_file(const char *fn, int t);
/* wraps caller-owned memory; destroy() frees only the context */
struct ioctx *ioctx_from_mem(const uint8_t *ptr, size_t sz);

#ifdef __cplusplus
}
#endif
#endif
