/* rqb_program.h -- the "solve program": what the host planner hands the device.
 *
 * All rows of a source block (received symbols, working rows, intermediate
 * symbols, emitted symbols) live in HBM, row-major, one pitch.  Row operations
 * are column-local (reference: deps/oblas/oblas_avx.c:62-73 works byte by byte),
 * so the block is cut into column slices (64, 128 or 256 bytes, chosen per
 * launch: the program does not depend on the width) and every CTA replays the
 * same program on its own slice with no inter-CTA traffic; its working set
 * (slice width x rows) is what stays hot in L1/L2.
 *
 * The program is a linear stream of fixed-size PAGES (staged into shared memory
 * by TMA bulk copies, see rqb_device.cu).  A page holds whole LEVELS; all tasks
 * of one level are independent, levels are separated by a CTA barrier.  A TASK
 * computes one destination row as a gather over at most RQB_MAX_SRCS source rows:
 *
 *     row[dst] = XOR_k  beta_k * row[src_k]
 *
 * which subsumes the reference's oaddrow/oaxpy/oscal sequences
 * (lib/precode.c:15-32) merged per destination, the final row permutation
 * (lib/precode.c:3-13,379-389: a copy into the C space) and the LT combine
 * (decode_row, lib/nanorq.c:184-204: a gather into the SYM space).
 * A destination that accumulates lists its own previous location as a source.
 *
 * All rows of a block sit in ONE arena, one pitch, in the order
 *     [ IN: in_rows | SYM: sym_rows | C: L | ZERO: 1 | WS: working rows ]
 * and a row reference is simply the 24-bit row number inside that arena, so the
 * kernel turns a reference into an address with one multiply-add.  The ZERO row
 * is never written: the source list of an XOR task is padded with it to exactly
 * 4 or 8 entries, so the kernel issues its loads without per-source predicates.
 */
#ifndef RQB_PROGRAM_H
#define RQB_PROGRAM_H

#include <stdint.h>

#define RQB_PAGE_BYTES 8192u /* multiple of 16 (TMA bulk copy granularity); 4 KiB pages were measured: more level
                                pieces and page turns, +15-25 % single-block latency in both kernels */
#define RQB_SLICE_BYTES 128u /* default column slice per CTA (rqb_device.cu picks 64 / 128 / 256 per launch) */
#define RQB_MAX_SRCS 8u      /* sources per XOR/GF task (the kernel keeps them all in flight) */
#define RQB_ROW_NONE 0xFFFFFFFFu

enum rqb_space { /* in arena order */
  RQB_SP_IN = 0,  /* received / source symbols as uploaded (read-only)      */
  RQB_SP_SYM = 1, /* emitted symbols (repair symbols / recovered symbols)   */
  RQB_SP_C = 2,   /* intermediate symbols C[0..L) in RFC order              */
  RQB_SP_WS = 3   /* working rows                                           */
};
#define RQB_MAX_ROWS 0x00FFFFFEu /* rows an arena can hold */
#define RQB_REF_MASK 0x00FFFFFFu
#define RQB_REF_NONE 0x00FFFFFFu /* SCAN: "no row here" */
/* GF source: row | beta << 24.  XOR and SCAN sources are the bare row number. */
#define RQB_SRC(ref, beta) ((uint32_t)(ref) | ((uint32_t)(beta) << 24))

enum rqb_task_kind {
  RQB_T_XOR = 0, /* row[dst] = XOR row[src_k]              nsrc <= RQB_MAX_SRCS (0 => zero row);
                    the list holds 4 (nsrc <= 4) or 8 entries, padded with the ZERO row            */
  RQB_T_GF = 1,  /* row[dst] = XOR beta_k * row[src_k]     nsrc <= RQB_MAX_SRCS                 */
  RQB_T_SCAN = 2, /* alpha-scan over nsrc entries, see below                                   */
  RQB_T_TAB = 3   /* table gather, see below                                                   */
};

/* TAB (the back-substitution x = Y ^ G z through 8-bit XOR tables, rqb_planner.c phase F):
 *     row[dst] = row[src0] ^ XOR_{j < nsrc, b_j != 0} row[tab_base + 256*j + b_j]
 * src0 is the task's `pad` word, the "source list" holds the nsrc bytes b_j (one row of the
 * bit matrix G, 8 inactive symbols per byte), tab_base and the ZERO row come from the level
 * header.  A third of the program bytes of three gather tasks per row. */

/* SCAN (restates the structure of precode_matrix_make_HDPC, lib/precode.c:60-83:
 * column j = alpha * column j+1 plus two ones, i.e. HDPC*C is a Horner scheme
 * in alpha).  For the entries e_0..e_{n-1}, starting from y = 0:
 *     y = alpha*y ^ row[ref(e_k)]      (ref NONE => just y = alpha*y)
 *     WS row idx(dst)+k = y
 * The HDPC rows are then ordinary XOR/GF gathers over the y rows. */

typedef struct {
  uint32_t src_off; /* byte offset of the source list from the page start (16-byte aligned) */
  uint32_t dst;     /* row number                                                           */
  uint16_t nsrc;
  uint8_t kind;
  uint8_t aux;
  uint32_t pad;     /* TAB: the row src0                                                    */
} rqb_task; /* 16 bytes: one 128-bit shared-memory load */

/* page := rqb_page_hdr, then levels back to back (each 16-byte aligned)
 * level := rqb_level_hdr, rqb_task[n_tasks], source lists (u32, each list 16-byte aligned) */
typedef struct {
  uint32_t n_levels;
  uint32_t pad[3];
} rqb_page_hdr;

typedef struct {
  uint32_t n_tasks;
  uint32_t next_off; /* byte offset (from page start) of the next level */
  uint32_t tab_base; /* first row of the XOR tables (TAB tasks)         */
  uint32_t zero_row; /* the all-zero row                                */
} rqb_level_hdr;

#define RQB_MAX_H 16u /* HDPC rows, RFC 6330 table 2: H <= 16 */

/* ---------------------------------------------------------------------------
 * The SHARED-MEMORY flavour of the program (rqb_solve_smem_kernel).
 *
 * For blocks whose rows fit, a CTA keeps its column slice (16, 32 or 64 bytes wide) of EVERY
 * live row in shared memory ("slots") for the whole solve: the received symbols are read from
 * HBM once (LOAD tasks), every row operation of the elimination runs in place on the slots,
 * and only the results (intermediate symbols / emitted symbols) are written back -- DRAM
 * traffic is the compulsory traffic.  Same page / level / task format; the differences:
 *
 *   row reference   bit 23 set   -> row (ref & 0x7FFFFF) of the block's HBM arena (IN, SYM, C);
 *                   bit 23 clear -> shared-memory slot `ref`.  Slot 0 is all zero and never
 *                   written (it pads XOR lists; level header zero_row = 0).
 *   XOR / GF        as above over mixed references; a destination that accumulates lists
 *                   itself as a source (rows are updated in place).  XOR: aux bit 0 = the
 *                   destination and every source are slots (no HBM reference).
 *   LOAD            slots dst .. dst+nsrc-1  =  arena rows pad .. pad+nsrc-1   (no list)
 *   SCAN2           alpha-scan with the HDPC sums folded in (the y_j are never stored):
 *                     y = alpha*y ^ row[ref(e_k)];  slot[dst + h1(e_k)] ^= y;  slot[dst + h2(e_k)] ^= y
 *                   entry e = ref | h1 << 24 | h2 << 28 (the two HDPC rows that have a one in
 *                   that column, lib/precode.c:68-81); the task first clears its H accumulator
 *                   slots dst..dst+H-1 (H = aux & 31) and finally stores y to slot `pad`.
 *                   aux bit 7: the last entry is the last column of the scan (no ones there).
 *   TAB             in place: slot[dst] ^= XOR_j slot[tab_base + (j << bits) + v_j], v_j != 0;
 *                   the list holds one byte v_j per group of `bits` = aux inactive symbols
 *                   (4..8, whatever the slot budget allows); pad = arena reference that also
 *                   receives the result (an encoder's row of the C space) or RQB_ROW_NONE.
 */
#define RQB_REF_GLOBAL 0x00800000u
#define RQB_T_LOAD 4
#define RQB_T_SCAN2 5
#define RQB_SMEM_RING_STAGES 3u
#define RQB_SMEM_BUDGET_BYTES (227u * 1024u - RQB_SMEM_RING_STAGES * RQB_PAGE_BYTES - 128u) /* slots of one CTA: 227 KB minus the page ring and its barriers */

#endif
