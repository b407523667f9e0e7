/* plan_interp.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU interpreter of the device solve program (nanorq_b200/csrc/
 * rqb_program.h).  It lets the CPU test-suite validate the host planner
 * against the oracle without a GPU, and it checks the property the kernel
 * relies on: within a level no task reads a slot another task writes.
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

/* returns 0 ok, 10 = intra-level hazard, 11 = malformed, 12 = misaligned (the kernel
 * reads level headers and tasks as 16-byte vectors, u16 source lists as 8-byte
 * vectors and u32 lists as words) */
int rqb_interp_run(uint32_t n_slots, const uint32_t *load_src, uint32_t n_pages, const uint8_t *pages,
                   const uint8_t *in, size_t in_pitch, size_t T, uint8_t *c_out, size_t c_pitch,
                   uint8_t *sym_out, size_t sym_pitch) {
  size_t ns = n_slots;
  uint8_t *ws = calloc(ns * T + 1, 1);
  uint32_t *wstamp = calloc(ns, sizeof(uint32_t)); /* level id that last wrote the slot */
  uint32_t *wowner = calloc(ns, sizeof(uint32_t));
  uint8_t *tmp = malloc(T ? T : 1);
  int rc = 0;
  uint32_t level_id = 0;
  for (size_t s = 0; s < ns; s++)
    if (load_src[s] != RQB_ROW_NONE) memcpy(ws + s * T, in + (size_t)load_src[s] * in_pitch, T);
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
      for (uint32_t k = 0; k < lh->n_tasks; k++) {
        const rqb_task *t = &tasks[k];
        if (t->kind == RQB_T_OUT_C || t->kind == RQB_T_OUT_SYM) continue;
        uint32_t cnt = t->kind == RQB_T_HORNER ? t->arg + 1 : 1;
        for (uint32_t q = 0; q < cnt; q++) {
          if (t->dst + q >= ns) { rc = 11; break; }
          if (wstamp[t->dst + q] == level_id) rc = 10; /* two writers */
          wstamp[t->dst + q] = level_id;
          wowner[t->dst + q] = k;
        }
      }
      /* pass 2: execute; sources must not be written in this level by another task */
      for (uint32_t k = 0; k < lh->n_tasks && !rc; k++) {
        const rqb_task *t = &tasks[k];
        const uint16_t *s16 = (const uint16_t *)(page + t->src_off);
        const uint32_t *s32 = (const uint32_t *)(page + t->src_off);
        if (t->src_off + (size_t)t->nsrc * 2 > RQB_PAGE_BYTES) { rc = 11; break; }
        if (t->src_off % 8) { rc = 12; break; }
        switch (t->kind) {
          case RQB_T_XOR_SET: case RQB_T_XOR_ACC: case RQB_T_LOAD_XOR:
          case RQB_T_OUT_C: case RQB_T_OUT_SYM: {
            memset(tmp, 0, T);
            if (t->kind == RQB_T_LOAD_XOR && t->arg != RQB_ROW_NONE)
              memcpy(tmp, in + (size_t)t->arg * in_pitch, T);
            for (uint32_t q = 0; q < t->nsrc; q++) {
              uint32_t s = s16[q];
              if (s >= ns) { rc = 11; break; }
              if (wstamp[s] == level_id && !(wowner[s] == k)) rc = 10;
              for (size_t b = 0; b < T; b++) tmp[b] ^= ws[s * T + b];
            }
            uint8_t *dst = t->kind == RQB_T_OUT_C ? c_out + (size_t)t->arg * c_pitch
                         : t->kind == RQB_T_OUT_SYM ? sym_out + (size_t)t->arg * sym_pitch
                         : ws + (size_t)t->dst * T;
            if (t->kind == RQB_T_XOR_ACC)
              for (size_t b = 0; b < T; b++) dst[b] ^= tmp[b];
            else
              memcpy(dst, tmp, T);
            break;
          }
          case RQB_T_GF_SET: case RQB_T_GF_ACC: {
            memset(tmp, 0, T);
            for (uint32_t q = 0; q < t->nsrc; q++) {
              uint32_t s = s32[q] & 0xFFFF;
              uint8_t beta = (uint8_t)(s32[q] >> 16);
              if (s >= ns) { rc = 11; break; }
              if (wstamp[s] == level_id && wowner[s] != k) rc = 10;
              for (size_t b = 0; b < T; b++) tmp[b] ^= gmul(ws[s * T + b], beta);
            }
            uint8_t *dst = ws + (size_t)t->dst * T;
            if (t->kind == RQB_T_GF_ACC)
              for (size_t b = 0; b < T; b++) dst[b] ^= tmp[b];
            else
              memcpy(dst, tmp, T);
            break;
          }
          case RQB_T_HORNER: {
            uint32_t Hh = t->arg;
            memset(ws + (size_t)t->dst * T, 0, (size_t)(Hh + 1) * T);
            memset(tmp, 0, T); /* tmp = y */
            for (uint32_t q = 0; q < t->nsrc; q++) {
              uint32_t e = s32[q], s = e & 0xFFFF, b1 = (e >> 16) & 15, b2 = (e >> 20) & 15;
              if (s != RQB_SLOT_NONE && s >= ns) { rc = 11; break; }
              if (s != RQB_SLOT_NONE && wstamp[s] == level_id) rc = 10;
              for (size_t b = 0; b < T; b++) {
                tmp[b] = gmul(tmp[b], 2);
                if (s != RQB_SLOT_NONE) tmp[b] ^= ws[s * T + b];
              }
              if (e >> 24 & 1)
                for (size_t b = 0; b < T; b++) {
                  ws[(size_t)(t->dst + b1) * T + b] ^= tmp[b];
                  ws[(size_t)(t->dst + b2) * T + b] ^= tmp[b];
                }
            }
            memcpy(ws + (size_t)(t->dst + Hh) * T, tmp, T);
            break;
          }
          default: rc = 11;
        }
      }
      off = lh->next_off;
    }
  }
  free(ws); free(wstamp); free(wowner); free(tmp);
  return rc;
}
