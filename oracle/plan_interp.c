/* plan_interp.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU interpreter of the device solve program (nanorq_b200/csrc/
 * rqb_program.h).  It lets the CPU test-suite validate the host planner
 * against the oracle without a GPU, and it checks the properties the kernel
 * relies on: alignment of the vector-loaded structures, at most RQB_MAX_SRCS
 * sources per XOR/GF task, and -- within a level -- that no task reads a row
 * another task writes and no two tasks write the same row.
 * It is written independently of the CUDA kernel and is never used by the
 * product path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rqb_program.h"

static uint8_t gmul(uint8_t a, uint8_t b) { /* shift-and-add, poly 0x11D */
  uint8_t r = 0;
  while (b) {
    if (b & 1) r ^= a;
    a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1D : 0));
    b >>= 1;
  }
  return r;
}

typedef struct {
  uint8_t *base[4];
  size_t pitch[4], rows[4];
  uint32_t *wstamp[4], *wowner[4]; /* level id / task that last wrote the row */
} spaces;

static uint8_t *row_ptr(spaces *sp, uint32_t ref, int *rc) {
  uint32_t s = (ref >> RQB_IDX_BITS) & 3u, idx = ref & (RQB_MAX_ROWS - 1);
  if (idx >= sp->rows[s]) {
    *rc = 11;
    return NULL;
  }
  return sp->base[s] + (size_t)idx * sp->pitch[s];
}

/* returns 0 ok, 10 = intra-level hazard, 11 = malformed, 12 = misaligned, 13 = too many sources,
 * 14 = a task writes the input space */
int rqb_interp_run(uint32_t n_ws_rows, uint32_t n_pages, const uint8_t *pages, const uint8_t *in, size_t in_rows,
                   size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows, size_t c_pitch, uint8_t *sym_out,
                   size_t sym_rows, size_t sym_pitch) {
  spaces sp;
  memset(&sp, 0, sizeof(sp));
  uint8_t *ws = malloc((size_t)n_ws_rows * T + 1);
  memset(ws, 0xA5, (size_t)n_ws_rows * T + 1); /* working rows start undefined, like device memory */
  sp.base[RQB_SP_IN] = (uint8_t *)in; sp.pitch[RQB_SP_IN] = in_pitch; sp.rows[RQB_SP_IN] = in_rows;
  sp.base[RQB_SP_WS] = ws; sp.pitch[RQB_SP_WS] = T; sp.rows[RQB_SP_WS] = n_ws_rows;
  sp.base[RQB_SP_C] = c_out; sp.pitch[RQB_SP_C] = c_pitch; sp.rows[RQB_SP_C] = c_rows;
  sp.base[RQB_SP_SYM] = sym_out; sp.pitch[RQB_SP_SYM] = sym_pitch; sp.rows[RQB_SP_SYM] = sym_rows;
  for (int s = 0; s < 4; s++) {
    sp.wstamp[s] = calloc(sp.rows[s] + 1, sizeof(uint32_t));
    sp.wowner[s] = calloc(sp.rows[s] + 1, sizeof(uint32_t));
  }
  uint8_t *tmp = malloc(T ? T : 1);
  int rc = 0;
  uint32_t level_id = 0;
#define STAMP(ref) sp.wstamp[((ref) >> RQB_IDX_BITS) & 3u][(ref) & (RQB_MAX_ROWS - 1)]
#define OWNER(ref) sp.wowner[((ref) >> RQB_IDX_BITS) & 3u][(ref) & (RQB_MAX_ROWS - 1)]
  for (uint32_t pg = 0; pg < n_pages && !rc; pg++) {
    const uint8_t *page = pages + (size_t)pg * RQB_PAGE_BYTES;
    const rqb_page_hdr *ph = (const rqb_page_hdr *)page;
    uint32_t off = sizeof(rqb_page_hdr);
    for (uint32_t lv = 0; lv < ph->n_levels && !rc; lv++) {
      if (off % 16 || off + sizeof(rqb_level_hdr) > RQB_PAGE_BYTES) { rc = off % 16 ? 12 : 11; break; }
      const rqb_level_hdr *lh = (const rqb_level_hdr *)(page + off);
      const rqb_task *tasks = (const rqb_task *)(page + off + sizeof(rqb_level_hdr));
      level_id++;
      /* pass 1: mark writers */
      for (uint32_t k = 0; k < lh->n_tasks && !rc; k++) {
        const rqb_task *t = &tasks[k];
        uint32_t cnt = t->kind == RQB_T_SCAN ? t->nsrc : 1u;
        if (((t->dst >> RQB_IDX_BITS) & 3u) == RQB_SP_IN) { rc = 14; break; }
        if (t->kind == RQB_T_SCAN && ((t->dst >> RQB_IDX_BITS) & 3u) != RQB_SP_WS) { rc = 11; break; }
        for (uint32_t q = 0; q < cnt; q++) {
          uint32_t ref = t->dst + q;
          if (!row_ptr(&sp, ref, &rc)) break;
          if (STAMP(ref) == level_id) rc = 10; /* two writers */
          STAMP(ref) = level_id;
          OWNER(ref) = k;
        }
      }
      /* pass 2: execute; sources must not be written in this level by another task */
      for (uint32_t k = 0; k < lh->n_tasks && !rc; k++) {
        const rqb_task *t = &tasks[k];
        const uint32_t *s32 = (const uint32_t *)(page + t->src_off);
        if (t->src_off + (size_t)t->nsrc * 4 > RQB_PAGE_BYTES) { rc = 11; break; }
        if (t->src_off % 16) { rc = 12; break; }
        switch (t->kind) {
          case RQB_T_XOR:
          case RQB_T_GF: {
            if (t->nsrc > RQB_MAX_SRCS) { rc = 13; break; }
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              uint32_t ref = s32[q] & RQB_REF_MASK;
              uint8_t beta = t->kind == RQB_T_GF ? (uint8_t)(s32[q] >> 24) : 1;
              const uint8_t *src = row_ptr(&sp, ref, &rc);
              if (!src) break;
              if (STAMP(ref) == level_id && OWNER(ref) != k) rc = 10;
              for (size_t b = 0; b < T; b++) tmp[b] ^= gmul(src[b], beta);
            }
            if (!rc) memcpy(row_ptr(&sp, t->dst, &rc), tmp, T);
            break;
          }
          case RQB_T_SCAN: {
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              uint32_t ref = s32[q] & RQB_REF_MASK;
              const uint8_t *src = NULL;
              if (ref != RQB_REF_NONE) {
                src = row_ptr(&sp, ref, &rc);
                if (!src) break;
                if (STAMP(ref) == level_id) rc = 10;
              }
              for (size_t b = 0; b < T; b++) tmp[b] = (uint8_t)(gmul(tmp[b], 2) ^ (src ? src[b] : 0));
              uint8_t *d = row_ptr(&sp, t->dst + q, &rc);
              if (d) memcpy(d, tmp, T);
            }
            break;
          }
          default: rc = 11;
        }
      }
      off = lh->next_off;
    }
  }
  for (int s = 0; s < 4; s++) {
    free(sp.wstamp[s]);
    free(sp.wowner[s]);
  }
  free(ws);
  free(tmp);
  return rc;
}
