/* rqb_gf256.h -- GF(2^8) host arithmetic, polynomial x^8+x^4+x^3+x^2+1 (0x11D),
 * generator alpha = 2.  Same field as the reference's generated tables
 * (deps/oblas/tablegen.c:10,31-52; octtables.h).  Header-only, host side. */
#ifndef RQB_GF256_H
#define RQB_GF256_H

#include <stdint.h>

typedef struct {
  uint8_t exp[512]; /* exp[e] = alpha^e, doubled so log(a)+log(b) needs no mod */
  uint8_t log[256]; /* log[0] unused (255) */
  uint8_t inv[256];
} rqb_gf_tables;

static inline void rqb_gf_build(rqb_gf_tables *t) {
  unsigned x = 1;
  for (int e = 0; e < 255; e++) {
    t->exp[e] = (uint8_t)x;
    t->log[x] = (uint8_t)e;
    x <<= 1;
    if (x & 0x100) x ^= 0x11D;
  }
  for (int e = 255; e < 512; e++) t->exp[e] = t->exp[e - 255];
  t->log[0] = 255;
  t->inv[0] = 0;
  for (int a = 1; a < 256; a++) t->inv[a] = t->exp[255 - t->log[a]];
}

static inline uint8_t rqb_gf_mul(const rqb_gf_tables *t, uint8_t a, uint8_t b) {
  return (a && b) ? t->exp[t->log[a] + t->log[b]] : 0;
}

/* alpha^e for any non-negative e */
static inline uint8_t rqb_gf_pow2(const rqb_gf_tables *t, long e) {
  return t->exp[e % 255];
}

#endif
