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
  uint8_t *arena;
  size_t pitch, rows;
  uint32_t *wstamp, *wowner; /* level id / task that last wrote the row */
} spaces;

static uint8_t *row_ptr(spaces *sp, uint32_t row, int *rc) {
  if (row >= sp->rows) {
    *rc = 11;
    return NULL;
  }
  return sp->arena + (size_t)row * sp->pitch;
}

/* Runs a program on an arena laid out as rqb_program.h says: row0[] = first row of
 * the spaces IN, SYM, C, WS; zero_row; n_rows in total.
 * returns 0 ok, 10 = intra-level hazard, 11 = malformed, 12 = misaligned, 13 = too many sources,
 * 14 = a task writes the input space or the ZERO row, 15 = an XOR list is not padded with the ZERO row */
int rqb_interp_run_ex(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                      const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                      size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch, int in_writable);

int rqb_interp_run(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                   const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                   size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch) {
  return rqb_interp_run_ex(row0, zero_row, n_rows, n_pages, pages, in, in_rows, in_pitch, T, c_out, c_rows, c_pitch,
                           sym_out, sym_rows, sym_pitch, 0);
}

/* in_writable: programs made from a reference schedule update the matrix rows (the
 * input space) in place; solve programs never write their inputs */
int rqb_interp_run_ex(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                      const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                      size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch, int in_writable) {
  spaces sp;
  memset(&sp, 0, sizeof(sp));
  if (row0[RQB_SP_IN] != 0 || row0[RQB_SP_SYM] < in_rows || row0[RQB_SP_C] < row0[RQB_SP_SYM] + sym_rows ||
      zero_row < row0[RQB_SP_C] + c_rows || row0[RQB_SP_WS] != zero_row + 1 || n_rows < row0[RQB_SP_WS])
    return 11;
  sp.rows = n_rows;
  sp.pitch = T;
  sp.arena = malloc((size_t)n_rows * T + 1);
  memset(sp.arena, 0xA5, (size_t)n_rows * T + 1); /* rows start undefined, like device memory */
  for (size_t r = 0; r < in_rows; r++) memcpy(sp.arena + r * T, in + r * in_pitch, T);
  memset(sp.arena + (size_t)zero_row * T, 0, T);
  sp.wstamp = calloc((size_t)n_rows + 1, sizeof(uint32_t));
  sp.wowner = calloc((size_t)n_rows + 1, sizeof(uint32_t));
  uint8_t *tmp = malloc(T ? T : 1);
  int rc = 0;
  uint32_t level_id = 0;
#define STAMP(ref) sp.wstamp[(ref)]
#define OWNER(ref) sp.wowner[(ref)]
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
        if ((!in_writable && t->dst < row0[RQB_SP_SYM]) || t->dst == zero_row) { rc = 14; break; }
        if (t->kind == RQB_T_SCAN && t->dst < row0[RQB_SP_WS]) { rc = 11; break; }
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
        if (t->kind != RQB_T_TAB && t->src_off + (size_t)t->nsrc * 4 > RQB_PAGE_BYTES) { rc = 11; break; }
        if (t->src_off % 16) { rc = 12; break; }
        switch (t->kind) {
          case RQB_T_XOR:
          case RQB_T_GF: {
            if (t->nsrc > RQB_MAX_SRCS) { rc = 13; break; }
            if (t->kind == RQB_T_XOR) { /* the kernel loads exactly 4 or 8 rows */
              uint32_t padded = t->nsrc <= 4 ? 4u : 8u;
              if (t->src_off + (size_t)padded * 4 > RQB_PAGE_BYTES) { rc = 11; break; }
              for (uint32_t q = t->nsrc; q < padded; q++)
                if (s32[q] != zero_row) rc = 15;
              if (rc) break;
            }
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              uint32_t ref = t->kind == RQB_T_GF ? (s32[q] & RQB_REF_MASK) : s32[q];
              uint8_t beta = t->kind == RQB_T_GF ? (uint8_t)(s32[q] >> 24) : 1;
              const uint8_t *src = row_ptr(&sp, ref, &rc);
              if (!src) break;
              if (STAMP(ref) == level_id && OWNER(ref) != k) rc = 10;
              for (size_t b = 0; b < T; b++) tmp[b] ^= gmul(src[b], beta);
            }
            if (!rc) memcpy(row_ptr(&sp, t->dst, &rc), tmp, T);
            break;
          }
          case RQB_T_TAB: { /* row[dst] = row[src0] ^ XOR_j table row (j, byte j), bytes where the list would be */
            if (t->src_off + (((size_t)t->nsrc + 15) & ~(size_t)15) > RQB_PAGE_BYTES) { rc = 11; break; }
            const uint8_t *bytes = (const uint8_t *)s32;
            const uint8_t *src = row_ptr(&sp, t->pad, &rc);
            if (!src) break;
            if (STAMP(t->pad) == level_id && OWNER(t->pad) != k) rc = 10;
            if (lh->zero_row != zero_row) rc = 11;
            memcpy(tmp, src, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              if (!bytes[q]) continue;
              uint32_t ref = lh->tab_base + 256u * q + bytes[q];
              const uint8_t *tr = row_ptr(&sp, ref, &rc);
              if (!tr) break;
              if (STAMP(ref) == level_id) rc = 10;
              for (size_t b = 0; b < T; b++) tmp[b] ^= tr[b];
            }
            for (uint32_t q = t->nsrc; q < ((t->nsrc + 15u) & ~15u); q++)
              if (bytes[q]) rc = 15; /* the kernel reads the list in 8-byte pieces: padding must be zero */
            if (!rc) memcpy(row_ptr(&sp, t->dst, &rc), tmp, T);
            break;
          }
          case RQB_T_SCAN: {
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              uint32_t ref = s32[q];
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
  for (size_t r = 0; r < c_rows; r++) memcpy(c_out + r * c_pitch, sp.arena + ((size_t)row0[RQB_SP_C] + r) * T, T);
  for (size_t r = 0; r < sym_rows; r++) memcpy(sym_out + r * sym_pitch, sp.arena + ((size_t)row0[RQB_SP_SYM] + r) * T, T);
  free(sp.wstamp);
  free(sp.wowner);
  free(sp.arena);
  free(tmp);
  return rc;
}
