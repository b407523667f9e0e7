/* rqb_io.c -- ioctx adapters (file / memory / mmap) with the behaviour of the
 * reference's lib/io.c:54,139,338.  Host I/O is outside the accelerated path;
 * the surface is kept so existing programs link. */
#define _DEFAULT_SOURCE
#define _FILE_OFFSET_BITS 64
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "io.h"
#include "nanorq_batch.h"
#include "rqb200.h"
#include "rqb_hostcopy.h"

/* ------------------------------------------------------------- stdio file */
typedef struct {
  struct ioctx io;
  FILE *fp;
} file_ctx;

static size_t f_read(struct ioctx *io, uint8_t *b, size_t n) { return fread(b, 1, n, ((file_ctx *)io)->fp); }
static size_t f_write(struct ioctx *io, const uint8_t *b, size_t n) { return fwrite(b, 1, n, ((file_ctx *)io)->fp); }
static bool f_seek(struct ioctx *io, const size_t off) { return fseeko(((file_ctx *)io)->fp, (off_t)off, SEEK_SET) == 0; }
static long f_tell(struct ioctx *io) { return ftell(((file_ctx *)io)->fp); }
static size_t f_size(struct ioctx *io) {
  FILE *fp = ((file_ctx *)io)->fp;
  off_t pos = ftello(fp);
  fseeko(fp, 0, SEEK_END);
  off_t end = ftello(fp);
  fseeko(fp, pos, SEEK_SET);
  return (size_t)end;
}
static void f_destroy(struct ioctx *io) {
  fclose(((file_ctx *)io)->fp);
  free(io);
}

struct ioctx *ioctx_from_file(const char *fn, int t) {
  FILE *fp = fopen(fn, t ? "r" : "w+");
  if (!fp) return NULL;
  file_ctx *c = calloc(1, sizeof(*c));
  c->fp = fp;
  c->io.read = f_read;
  c->io.write = f_write;
  c->io.seek = f_seek;
  c->io.size = f_size;
  c->io.tell = f_tell;
  c->io.destroy = f_destroy;
  c->io.seekable = true;
  c->io.writable = (t == 0);
  return &c->io;
}

/* ----------------------------------------------------------------- memory */
typedef struct {
  struct ioctx io;
  uint8_t *base;
  size_t pos, len;
  int pinned;  /* 1: page-locked by the caller (rqb_host_alloc / rqb_host_pin), 2: page-locked here */
} mem_ctx;

static size_t m_clip(mem_ctx *c, size_t n) { return c->pos + n > c->len ? c->len - c->pos : n; }
static size_t m_read(struct ioctx *io, uint8_t *b, size_t n) {
  mem_ctx *c = (mem_ctx *)io;
  n = m_clip(c, n);
  rqb_copy_stream(b, c->base + c->pos, n); /* the library reads into pinned staging rows */
  c->pos += n;
  return n;
}
static size_t m_write(struct ioctx *io, const uint8_t *b, size_t n) {
  mem_ctx *c = (mem_ctx *)io;
  n = m_clip(c, n);
  rqb_copy_stream(c->base + c->pos, b, n); /* decoded output: not read again soon */
  c->pos += n;
  return n;
}
static bool m_seek(struct ioctx *io, const size_t off) {
  mem_ctx *c = (mem_ctx *)io;
  if (off >= c->len) return false; /* like the reference: seeking to the end fails */
  c->pos = off;
  return true;
}
static long m_tell(struct ioctx *io) { return (long)((mem_ctx *)io)->pos; }
static size_t m_size(struct ioctx *io) { return ((mem_ctx *)io)->len; }
static void m_destroy(struct ioctx *io) {
  rqb_copy_fence(); /* rows were written with non-temporal stores */
  if (((mem_ctx *)io)->pinned == 2) rqb_host_unpin(((mem_ctx *)io)->base);
  free(io);
}

struct ioctx *ioctx_from_mem(const uint8_t *ptr, size_t sz) {
  mem_ctx *c = calloc(1, sizeof(*c));
  if (!c) return NULL;
  c->base = (uint8_t *)ptr;
  c->len = sz;
  c->io.read = m_read;
  c->io.write = m_write;
  c->io.seek = m_seek;
  c->io.size = m_size;
  c->io.tell = m_tell;
  c->io.destroy = m_destroy;
  c->io.seekable = true;
  c->io.writable = true;
  return &c->io;
}

/* Memory the GPU's copy engines can reach directly (the pinned/registered-memory ioctx of SURVEY
 * 8(f)4; the reference's ioctx_from_mem, lib/io.c:139-157, always goes through memcpy): same
 * behaviour as ioctx_from_mem for every caller of the vtable, but the nanorq_* calls of this
 * library recognise it (rqb_ioctx_mem_view) and move symbols between this memory and the device by
 * DMA, without the copy through a staging row.
 * already_pinned != 0: the caller page-locked the memory (rqb_host_alloc, rqb_host_pin, or
 * cudaHostAlloc / cudaHostRegister of its own); 0: it is page-locked here for the lifetime of
 * the context (takes ~0.2 ms per MiB: meant for long-lived buffers). */
struct ioctx *ioctx_from_pinned_mem(uint8_t *ptr, size_t sz, int already_pinned) {
  if (!ptr || !sz) return NULL;
  int pinned = 1;
  if (!already_pinned) {
    if (rqb_host_pin(ptr, sz) != 0) return NULL;
    pinned = 2;
  }
  struct ioctx *io = ioctx_from_mem(ptr, sz);
  if (!io) {
    if (pinned == 2) rqb_host_unpin(ptr);
    return NULL;
  }
  ((mem_ctx *)io)->pinned = pinned;
  return io;
}

/* library-internal: is `io` one of the memory contexts above?  Gives its span. */
int rqb_ioctx_mem_view(struct ioctx *io, uint8_t **base, size_t *len, int *pinned) {
  if (!io || io->read != m_read || io->destroy != m_destroy) return 0;
  const mem_ctx *c = (const mem_ctx *)io;
  *base = c->base;
  *len = c->len;
  *pinned = c->pinned != 0;
  return 1;
}

/* ------------------------------------------------------------------- mmap
 * Whole-file mapping.  A decoder-side file (t = 0) starts empty and grows on
 * demand; it is truncated to the highest byte written when destroyed. */
typedef struct {
  struct ioctx io;
  int fd;
  uint8_t *map;
  size_t maplen, pos, hiwater;
  bool writable;
} mmap_ctx;

static bool mm_grow(mmap_ctx *c, size_t need) {
  if (need <= c->maplen) return true;
  if (!c->writable) return false;
  size_t nl = c->maplen ? c->maplen : (size_t)1 << 16;
  while (nl < need) nl *= 2;
  if (ftruncate(c->fd, (off_t)nl) != 0) return false;
  if (c->map) munmap(c->map, c->maplen);
  c->map = mmap(NULL, nl, PROT_READ | PROT_WRITE, MAP_SHARED, c->fd, 0);
  if (c->map == MAP_FAILED) {
    c->map = NULL;
    c->maplen = 0;
    return false;
  }
  c->maplen = nl;
  return true;
}
static size_t mm_read(struct ioctx *io, uint8_t *b, size_t n) {
  mmap_ctx *c = (mmap_ctx *)io;
  size_t lim = c->writable ? c->hiwater : c->maplen;
  if (c->pos >= lim) return 0;
  if (c->pos + n > lim) n = lim - c->pos;
  memcpy(b, c->map + c->pos, n);
  c->pos += n;
  return n;
}
static size_t mm_write(struct ioctx *io, const uint8_t *b, size_t n) {
  mmap_ctx *c = (mmap_ctx *)io;
  if (!c->writable || !mm_grow(c, c->pos + n)) return 0;
  memcpy(c->map + c->pos, b, n);
  c->pos += n;
  if (c->pos > c->hiwater) c->hiwater = c->pos;
  return n;
}
static bool mm_seek(struct ioctx *io, const size_t off) {
  mmap_ctx *c = (mmap_ctx *)io;
  if (!c->writable && off >= c->maplen) return false;
  c->pos = off;
  return true;
}
static long mm_tell(struct ioctx *io) { return (long)((mmap_ctx *)io)->pos; }
static size_t mm_size(struct ioctx *io) {
  mmap_ctx *c = (mmap_ctx *)io;
  return c->writable ? c->hiwater : c->maplen;
}
static void mm_destroy(struct ioctx *io) {
  mmap_ctx *c = (mmap_ctx *)io;
  if (c->map) munmap(c->map, c->maplen);
  if (c->writable && ftruncate(c->fd, (off_t)c->hiwater) != 0) { /* best effort */ }
  close(c->fd);
  free(c);
}

struct ioctx *ioctx_mmap_file(const char *fn, int t) {
  int fd = t ? open(fn, O_RDONLY) : open(fn, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return NULL;
  mmap_ctx *c = calloc(1, sizeof(*c));
  c->fd = fd;
  c->writable = (t == 0);
  if (t) {
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) {
      close(fd);
      free(c);
      return NULL;
    }
    c->maplen = (size_t)st.st_size;
    c->map = mmap(NULL, c->maplen, PROT_READ, MAP_PRIVATE, fd, 0);
    if (c->map == MAP_FAILED) {
      close(fd);
      free(c);
      return NULL;
    }
  }
  c->io.read = mm_read;
  c->io.write = mm_write;
  c->io.seek = mm_seek;
  c->io.size = mm_size;
  c->io.tell = mm_tell;
  c->io.destroy = mm_destroy;
  c->io.seekable = true;
  c->io.writable = c->writable;
  return &c->io;
}
