/* plan_interp.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU interpreter of the device solve program (nanorq_b200/csrc/
 * rqb_program.h), both flavours.  It lets the CPU test-suite validate the host
 * planner against the oracle without a GPU, and it checks the properties the
 * kernels rely on: alignment of the vector-loaded structures, at most
 * RQB_MAX_SRCS sources per XOR/GF task, XOR lists padded with the zero row, and
 * -- within a level -- that no task reads a row another task writes and no two
 * tasks write the same row.  Rows start out as garbage (0xA5), like device
 * memory, so a program that reads a row nobody wrote does not pass.
 * It is written independently of the CUDA kernels and is never used by the
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
  uint8_t *store;            /* arena rows, then (smem flavour) the shared-memory slots */
  size_t pitch, rows;        /* rows of the store */
  size_t arena_rows;         /* rows of the HBM arena proper */
  int smem;                  /* smem flavour: a reference is a slot unless RQB_REF_GLOBAL is set */
  uint32_t zero_index;       /* store index of the all-zero row (arena ZERO row / slot 0) */
  uint32_t *wstamp, *wowner; /* level id / task that last wrote the row */
  uint32_t level_id;
} spaces;

#define BAD 0xFFFFFFFFu
/* reference -> index into the store; BAD = malformed */
static uint32_t ref_index(const spaces *sp, uint32_t ref) {
  if (!sp->smem) return ref < sp->arena_rows ? ref : BAD;
  if (ref & RQB_REF_GLOBAL) {
    ref &= RQB_REF_GLOBAL - 1u;
    return ref < sp->arena_rows ? ref : BAD;
  }
  return sp->arena_rows + ref < sp->rows ? (uint32_t)sp->arena_rows + ref : BAD;
}
static uint8_t *at(spaces *sp, uint32_t index) { return sp->store + (size_t)index * sp->pitch; }

/* a source row: must exist and must not be written in this level by another task */
static const uint8_t *src_row(spaces *sp, uint32_t ref, uint32_t task, int *rc) {
  uint32_t ix = ref_index(sp, ref);
  if (ix == BAD) {
    *rc = 11;
    return NULL;
  }
  if (sp->wstamp[ix] == sp->level_id && sp->wowner[ix] != task) *rc = 10;
  return at(sp, ix);
}

/* mark `ref` as written by `task` in this level: 10 = second writer, 14 = protected row */
static void mark_write(spaces *sp, uint32_t ref, uint32_t task, uint32_t protect_below, int *rc) {
  uint32_t ix = ref_index(sp, ref);
  if (ix == BAD) {
    *rc = 11;
    return;
  }
  if (ix == sp->zero_index || (ix < sp->arena_rows && ix < protect_below)) {
    *rc = 14;
    return;
  }
  if (sp->wstamp[ix] == sp->level_id) *rc = 10;
  sp->wstamp[ix] = sp->level_id;
  sp->wowner[ix] = task;
}

/* Runs a program on an arena laid out as rqb_program.h says: row0[] = first row of
 * the spaces IN, SYM, C, WS; zero_row; n_rows in total; smem != 0: the shared-memory
 * flavour with n_slots slots.
 * returns 0 ok, 10 = intra-level hazard, 11 = malformed, 12 = misaligned, 13 = too many sources,
 * 14 = a task writes the input space or the zero row, 15 = an XOR list is not padded with the zero row */
int rqb_interp_run2(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                    const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                    size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch, int in_writable, int smem,
                    uint32_t n_slots) {
  spaces sp;
  memset(&sp, 0, sizeof(sp));
  if (row0[RQB_SP_IN] != 0 || row0[RQB_SP_SYM] < in_rows || row0[RQB_SP_C] < row0[RQB_SP_SYM] + sym_rows ||
      zero_row < row0[RQB_SP_C] + c_rows || row0[RQB_SP_WS] != zero_row + 1 || n_rows < row0[RQB_SP_WS])
    return 11;
  if (smem && n_slots == 0) return 11;
  sp.smem = smem;
  sp.arena_rows = n_rows;
  sp.rows = (size_t)n_rows + (smem ? n_slots : 0);
  sp.pitch = T;
  sp.store = malloc(sp.rows * T + 1);
  memset(sp.store, 0xA5, sp.rows * T + 1); /* rows start undefined, like device and shared memory */
  for (size_t r = 0; r < in_rows; r++) memcpy(sp.store + r * T, in + r * in_pitch, T);
  memset(sp.store + (size_t)zero_row * T, 0, T);
  sp.zero_index = smem ? n_rows : zero_row; /* slot 0 */
  memset(at(&sp, sp.zero_index), 0, T);
  sp.wstamp = calloc(sp.rows + 1, sizeof(uint32_t));
  sp.wowner = calloc(sp.rows + 1, sizeof(uint32_t));
  uint8_t *tmp = malloc(T ? T : 1);
  const uint32_t protect = in_writable ? 0u : row0[RQB_SP_SYM]; /* solve programs never write their inputs */
  const uint32_t list_zero = smem ? 0u : zero_row;
  int rc = 0;
  for (uint32_t pg = 0; pg < n_pages && !rc; pg++) {
    const uint8_t *page = pages + (size_t)pg * RQB_PAGE_BYTES;
    const rqb_page_hdr *ph = (const rqb_page_hdr *)page;
    uint32_t off = sizeof(rqb_page_hdr);
    for (uint32_t lv = 0; lv < ph->n_levels && !rc; lv++) {
      if (off % 16 || off + sizeof(rqb_level_hdr) > RQB_PAGE_BYTES) { rc = off % 16 ? 12 : 11; break; }
      const rqb_level_hdr *lh = (const rqb_level_hdr *)(page + off);
      const rqb_task *tasks = (const rqb_task *)(page + off + sizeof(rqb_level_hdr));
      sp.level_id++;
      if (lh->zero_row != list_zero) rc = 11;
      /* pass 1: mark writers */
      for (uint32_t k = 0; k < lh->n_tasks && !rc; k++) {
        const rqb_task *t = &tasks[k];
        switch (t->kind) {
          case RQB_T_SCAN:
            if (smem || t->dst < row0[RQB_SP_WS]) { rc = 11; break; }
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) mark_write(&sp, t->dst + q, k, protect, &rc);
            break;
          case RQB_T_LOAD:
            if (!smem) { rc = 11; break; }
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) mark_write(&sp, t->dst + q, k, protect, &rc);
            break;
          case RQB_T_SCAN2: {
            if (!smem) { rc = 11; break; }
            const uint32_t H = t->aux & 31u;
            for (uint32_t q = 0; q < H && !rc; q++) mark_write(&sp, t->dst + q, k, protect, &rc);
            if (!rc) mark_write(&sp, t->pad, k, protect, &rc);
            break;
          }
          case RQB_T_TAB:
            mark_write(&sp, t->dst, k, protect, &rc);
            if (smem && t->pad != RQB_ROW_NONE && !rc) mark_write(&sp, t->pad, k, protect, &rc);
            break;
          default:
            mark_write(&sp, t->dst, k, protect, &rc);
        }
      }
      /* pass 2: execute; sources must not be written in this level by another task */
      for (uint32_t k = 0; k < lh->n_tasks && !rc; k++) {
        const rqb_task *t = &tasks[k];
        const uint32_t *s32 = (const uint32_t *)(page + t->src_off);
        if (t->kind != RQB_T_LOAD) {
          if (t->kind != RQB_T_TAB && t->src_off + (size_t)t->nsrc * 4 > RQB_PAGE_BYTES) { rc = 11; break; }
          if (t->src_off % 16) { rc = 12; break; }
        }
        switch (t->kind) {
          case RQB_T_XOR:
          case RQB_T_GF: {
            if (t->nsrc > RQB_MAX_SRCS) { rc = 13; break; }
            if (t->kind == RQB_T_XOR) { /* the kernels load exactly 4 or 8 rows */
              uint32_t padded = t->nsrc <= 4 ? 4u : 8u;
              if (t->src_off + (size_t)padded * 4 > RQB_PAGE_BYTES) { rc = 11; break; }
              for (uint32_t q = t->nsrc; q < padded; q++)
                if (s32[q] != list_zero) rc = 15;
              if (smem && (t->aux & 1u)) { /* the kernel's lean path: no reference may be an HBM row */
                if (t->dst & RQB_REF_GLOBAL) rc = 11;
                for (uint32_t q = 0; q < t->nsrc; q++)
                  if (s32[q] & RQB_REF_GLOBAL) rc = 11;
              }
              if (rc) break;
            }
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              uint32_t ref = t->kind == RQB_T_GF ? (s32[q] & RQB_REF_MASK) : s32[q];
              uint8_t beta = t->kind == RQB_T_GF ? (uint8_t)(s32[q] >> 24) : 1;
              const uint8_t *src = src_row(&sp, ref, k, &rc);
              if (!src) break;
              for (size_t b = 0; b < T; b++) tmp[b] ^= gmul(src[b], beta);
            }
            if (!rc) memcpy(at(&sp, ref_index(&sp, t->dst)), tmp, T);
            break;
          }
          case RQB_T_TAB: {
            /* HBM flavour: row[dst] = row[pad] ^ XOR_j table row (256*j + byte j)
             * smem flavour: slot[dst] ^= XOR_j slot (j << aux) + byte j, result also to arena row pad */
            if (t->src_off + (((size_t)t->nsrc + 15) & ~(size_t)15) > RQB_PAGE_BYTES) { rc = 11; break; }
            const uint8_t *bytes = (const uint8_t *)s32;
            const uint32_t bits = smem ? t->aux : 8u;
            if (smem && (bits < 4 || bits > 8)) { rc = 11; break; }
            const uint8_t *src = src_row(&sp, smem ? t->dst : t->pad, k, &rc);
            if (!src) break;
            memcpy(tmp, src, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              if (!bytes[q]) continue;
              if (bytes[q] >> bits) { rc = 11; break; }
              const uint8_t *tr = src_row(&sp, lh->tab_base + (q << bits) + bytes[q], 0xFFFFFFFFu, &rc);
              if (!tr) break;
              for (size_t b = 0; b < T; b++) tmp[b] ^= tr[b];
            }
            for (uint32_t q = t->nsrc; q < ((t->nsrc + 15u) & ~15u); q++)
              if (bytes[q]) rc = 15; /* the kernels read the list in 8-byte pieces: padding must be zero */
            if (rc) break;
            memcpy(at(&sp, ref_index(&sp, t->dst)), tmp, T);
            if (smem && t->pad != RQB_ROW_NONE) memcpy(at(&sp, ref_index(&sp, t->pad)), tmp, T);
            break;
          }
          case RQB_T_SCAN: {
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              const uint8_t *src = NULL;
              if (s32[q] != RQB_REF_NONE) {
                src = src_row(&sp, s32[q], 0xFFFFFFFFu, &rc);
                if (!src) break;
              }
              for (size_t b = 0; b < T; b++) tmp[b] = (uint8_t)(gmul(tmp[b], 2) ^ (src ? src[b] : 0));
              memcpy(at(&sp, ref_index(&sp, t->dst + q)), tmp, T);
            }
            break;
          }
          case RQB_T_LOAD: {
            if (!(t->pad & RQB_REF_GLOBAL)) { rc = 11; break; }
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              const uint8_t *src = src_row(&sp, t->pad + q, k, &rc);
              if (!src) break;
              memcpy(at(&sp, ref_index(&sp, t->dst + q)), src, T);
            }
            break;
          }
          case RQB_T_SCAN2: {
            const uint32_t H = t->aux & 31u, last_plain = t->aux & 0x80u;
            if (H == 0 || H > RQB_MAX_H) { rc = 11; break; }
            for (uint32_t h = 0; h < H; h++) memset(at(&sp, ref_index(&sp, t->dst + h)), 0, T);
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc && !rc; q++) {
              const uint32_t ref = s32[q] & RQB_REF_MASK, h1 = (s32[q] >> 24) & 15u, h2 = s32[q] >> 28;
              const uint8_t *src = NULL;
              if (ref != RQB_REF_NONE) {
                if (ref & RQB_REF_GLOBAL) { rc = 11; break; } /* scans run on slots only */
                src = src_row(&sp, ref, 0xFFFFFFFFu, &rc);
                if (!src) break;
              }
              for (size_t b = 0; b < T; b++) tmp[b] = (uint8_t)(gmul(tmp[b], 2) ^ (src ? src[b] : 0));
              if (last_plain && q + 1 == t->nsrc) continue;
              if (h1 >= H || h2 >= H || h1 == h2) { rc = 11; break; }
              uint8_t *a1 = at(&sp, ref_index(&sp, t->dst + h1)), *a2 = at(&sp, ref_index(&sp, t->dst + h2));
              for (size_t b = 0; b < T; b++) {
                a1[b] ^= tmp[b];
                a2[b] ^= tmp[b];
              }
            }
            if (!rc) memcpy(at(&sp, ref_index(&sp, t->pad)), tmp, T);
            break;
          }
          default: rc = 11;
        }
      }
      off = lh->next_off;
    }
  }
  for (size_t r = 0; r < c_rows; r++) memcpy(c_out + r * c_pitch, sp.store + ((size_t)row0[RQB_SP_C] + r) * T, T);
  for (size_t r = 0; r < sym_rows; r++) memcpy(sym_out + r * sym_pitch, sp.store + ((size_t)row0[RQB_SP_SYM] + r) * T, T);
  free(sp.wstamp);
  free(sp.wowner);
  free(sp.store);
  free(tmp);
  return rc;
}

/* in_writable: programs made from a reference schedule update the matrix rows (the
 * input space) in place; solve programs never write their inputs */
int rqb_interp_run_ex(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                      const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                      size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch, int in_writable) {
  return rqb_interp_run2(row0, zero_row, n_rows, n_pages, pages, in, in_rows, in_pitch, T, c_out, c_rows, c_pitch,
                         sym_out, sym_rows, sym_pitch, in_writable, 0, 0);
}

int rqb_interp_run(const uint32_t *row0, uint32_t zero_row, uint32_t n_rows, uint32_t n_pages, const uint8_t *pages,
                   const uint8_t *in, size_t in_rows, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_rows,
                   size_t c_pitch, uint8_t *sym_out, size_t sym_rows, size_t sym_pitch) {
  return rqb_interp_run_ex(row0, zero_row, n_rows, n_pages, pages, in, in_rows, in_pitch, T, c_out, c_rows, c_pitch,
                           sym_out, sym_rows, sym_pitch, 0);
}
