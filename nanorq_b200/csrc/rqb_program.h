/* rqb_program.h -- the "solve program": what the host planner hands the device.
 *
 * A source block's symbol matrix is processed by CTAs that each own a w-byte
 * column slice of EVERY row in shared memory ("slots").  Row operations are
 * column-local (reference: deps/oblas/oblas_avx.c:62-73 works byte by byte), so
 * every CTA replays the same program on its own slice with no inter-CTA traffic.
 *
 * The program is a linear stream of fixed-size PAGES (staged into shared memory
 * by TMA bulk copies, see rqb_device.cu).  A page holds whole LEVELS; all tasks
 * of one level are independent, levels are separated by a CTA barrier.  A TASK
 * computes one destination row as a gather over source slots:
 *
 *     dst (=|^=)  [in[arg]] ^ XOR_k  beta_k * ws[src_k]
 *
 * which subsumes the reference's oaddrow/oaxpy/oscal sequences
 * (lib/precode.c:15-32) merged per destination, the final row permutation
 * (lib/precode.c:3-13,379-389: expressed as OUT tasks) and the LT combine
 * (decode_row, lib/nanorq.c:184-204: an OUT task with several sources).
 */
#ifndef RQB_PROGRAM_H
#define RQB_PROGRAM_H

#include <stdint.h>

#define RQB_PAGE_BYTES 8192u /* multiple of 16 (TMA bulk copy granularity) */
#define RQB_SLOT_NONE 0xFFFFu
#define RQB_ROW_NONE 0xFFFFFFFFu
#define RQB_MAX_SLOTS 65535u

enum rqb_task_kind {
  RQB_T_XOR_SET = 0,  /* ws[dst]  = XOR ws[src]                      srcs: u16          */
  RQB_T_XOR_ACC = 1,  /* ws[dst] ^= XOR ws[src]                      srcs: u16          */
  RQB_T_GF_SET = 2,   /* ws[dst]  = XOR beta*ws[src]                 srcs: u32 slot|beta<<16 */
  RQB_T_GF_ACC = 3,   /* ws[dst] ^= XOR beta*ws[src]                 srcs: u32          */
  RQB_T_LOAD_XOR = 4, /* ws[dst]  = in[arg] ^ XOR ws[src]  (arg may be RQB_ROW_NONE)   */
  RQB_T_OUT_C = 5,    /* c_out[arg]   = XOR ws[src]                  srcs: u16          */
  RQB_T_OUT_SYM = 6,  /* sym_out[arg] = XOR ws[src]                  srcs: u16          */
  RQB_T_HORNER = 7    /* HDPC chunk scan, see below                  srcs: u32          */
};

/* HORNER (restates the structure of precode_matrix_make_HDPC, lib/precode.c:60-83:
 * column j = alpha * column j+1 plus two ones).  For the entries e_0..e_{n-1}:
 *     y = alpha*y ^ ws[slot(e)]   (slot NONE => just y = alpha*y)
 *     if (flag(e)) { acc[b1(e)] ^= y; acc[b2(e)] ^= y; }
 * acc[h] lives in ws[dst+h], h < arg (=H); the final y goes to ws[dst+arg].
 * entry = slot | b1<<16 | b2<<20 | flag<<24. */
#define RQB_HORNER_ENTRY(slot, b1, b2, flag) \
  ((uint32_t)(slot) | ((uint32_t)(b1) << 16) | ((uint32_t)(b2) << 20) | ((uint32_t)(flag) << 24))

typedef struct {
  uint32_t src_off; /* byte offset of the source list from the page start */
  uint32_t arg;     /* in row / out row / H, by kind                      */
  uint16_t nsrc;
  uint16_t dst;
  uint8_t kind;
  uint8_t pad[3];
} rqb_task; /* 16 bytes: one 128-bit shared-memory load */

/* page := rqb_page_hdr, then levels back to back (each 16-byte aligned)
 * level := rqb_level_hdr, rqb_task[n_tasks], source lists                */
typedef struct {
  uint32_t n_levels;
  uint32_t pad[3];
} rqb_page_hdr;

typedef struct {
  uint32_t n_tasks;
  uint32_t next_off; /* byte offset (from page start) of the next level */
  uint32_t pad[2];
} rqb_level_hdr;

#endif
