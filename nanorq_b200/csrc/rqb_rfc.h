/* rqb_rfc.h -- RFC 6330 code-construction arithmetic shared by the host planner
 * (C) and the device LT kernel (CUDA).  Header-only so the same integer math is
 * compiled for both sides; the tables are passed in (host: rfc6330_tables.h,
 * device: a __constant__ copy).
 *
 * Behavioural reference: lib/params.c:21-65, lib/tuple.c:13-43, lib/rand.c:183-190.
 */
#ifndef RQB_RFC_H
#define RQB_RFC_H

#include <stdint.h>

#ifdef __CUDACC__
#define RQB_HD __host__ __device__ __forceinline__
#else
#define RQB_HD static inline
#endif

typedef struct {
  int Kprime, S, H, W, L, P, P1, U, B, J;
} rqb_params;

typedef struct {
  uint32_t d, a, b, d1, a1, b1;
} rqb_tuple;

#define RQB_MAX_LT_DEGREE 40 /* d <= 30, d1 <= 3 */

/* Rand[y, i, m]  (RFC 6330 5.3.5.1; lib/rand.c:183-190) */
RQB_HD uint32_t rqb_rand(const uint32_t (*V)[256], uint32_t y, uint32_t i, uint32_t m) {
  uint32_t r = V[0][(y + i) & 0xff] ^ V[1][((y >> 8) + i) & 0xff] ^
               V[2][((y >> 16) + i) & 0xff] ^ V[3][((y >> 24) + i) & 0xff];
  return r % m;
}

/* Deg[v]  (RFC 6330 5.3.5.2; lib/tuple.c:13-19) */
RQB_HD uint32_t rqb_degree(const uint32_t *cdf, uint32_t v, uint32_t W) {
  uint32_t d = 0;
  while (d < 30 && v >= cdf[d]) d++;
  return d < W - 2 ? d : W - 2;
}

/* Tuple[K', X]  (RFC 6330 5.3.5.4; lib/tuple.c:21-43).  The multiply-add wraps
 * modulo 2^32 exactly like the reference's (uint32_t) cast. */
RQB_HD rqb_tuple rqb_tuple_gen(const rqb_params *P, const uint32_t (*V)[256],
                               const uint32_t *cdf, uint32_t X) {
  rqb_tuple t;
  uint32_t A = 53591u + (uint32_t)P->J * 997u;
  A |= 1u;
  uint32_t Bv = 10267u * ((uint32_t)P->J + 1u);
  uint32_t y = Bv + X * A;
  uint32_t v = rqb_rand(V, y, 0, 1u << 20);
  t.d = rqb_degree(cdf, v, (uint32_t)P->W);
  t.a = 1u + rqb_rand(V, y, 1, (uint32_t)P->W - 1u);
  t.b = rqb_rand(V, y, 2, (uint32_t)P->W);
  t.d1 = (t.d < 4) ? 2u + rqb_rand(V, X, 3, 2) : 2u;
  t.a1 = 1u + rqb_rand(V, X, 4, (uint32_t)P->P1 - 1u);
  t.b1 = rqb_rand(V, X, 5, (uint32_t)P->P1);
  return t;
}

/* The intermediate-symbol indices an encoding symbol with internal id X is the
 * XOR of (RFC 6330 5.3.5.3; lib/params.c:47-65).  Returns the count. */
RQB_HD int rqb_lt_indices(const rqb_params *P, const uint32_t (*V)[256],
                          const uint32_t *cdf, uint32_t X, uint32_t *out) {
  rqb_tuple t = rqb_tuple_gen(P, V, cdf, X);
  const uint32_t W = (uint32_t)P->W, Pn = (uint32_t)P->P, P1 = (uint32_t)P->P1;
  int n = 0;
  uint32_t b = t.b;
  out[n++] = b;
  for (uint32_t j = 1; j < t.d; j++) {
    b += t.a;
    if (b >= W) b -= W;
    out[n++] = b;
  }
  uint32_t b1 = t.b1;
  while (b1 >= Pn) b1 = (b1 + t.a1) % P1;
  out[n++] = W + b1;
  for (uint32_t j = 1; j < t.d1; j++) {
    do {
      b1 = (b1 + t.a1) % P1;
    } while (b1 >= Pn);
    out[n++] = W + b1;
  }
  return n;
}

#endif
