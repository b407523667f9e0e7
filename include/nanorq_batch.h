/* nanorq_batch.h -- batch entry points next to the reference's per-symbol API.
 *
 * nanorq.h (the drop-in surface) moves one symbol per call: nanorq_encode
 * (reference lib/nanorq.c:403-435) and nanorq_decoder_add_symbol (:478-509) each
 * cost a call and a host copy per T-byte symbol, and with the solve on the GPU those
 * copies are what a round trip spends its time on.  The calls
 * below move a whole range of symbols per call, and when the caller's buffers are
 * page-locked (rqb_host_alloc / rqb_host_pin in rqb200.h, ioctx_from_pinned_mem
 * here) the symbols travel between those buffers and the device by DMA only -- no
 * CPU copy, no staging row.  Same objects, same OTI, same bytes as the per-symbol
 * calls; the two families can be mixed on one object.
 */
#ifndef NANORQ_BATCH_H
#define NANORQ_BATCH_H

#include "nanorq.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ioctx over page-locked memory (cf. ioctx_from_mem, lib/io.c:139-157) */
struct ioctx *ioctx_from_pinned_mem(uint8_t *ptr, size_t sz, int already_pinned);

/* ---- encoder: symbols esi0 .. esi0+n-1 of block sbn into dst, row k at dst + k*pitch
 * (T bytes each; pitch >= T).  Source and repair ESIs may be mixed in one range.
 * Replaces n calls of nanorq_encode (lib/nanorq.c:403-435); generates the block's
 * intermediate symbols first if that has not happened yet (nanorq_generate_symbols).
 * Returns the number of symbols written: n, or 0 on failure. */
size_t nanorq_encode_range(nanorq *rq, uint8_t sbn, uint32_t esi0, uint32_t n, void *dst, size_t pitch,
                           struct ioctx *io);

/* a row of the caller's buffer that holds no symbol: a receive ring indexed by sequence number
 * has such holes where packets were lost; the row is skipped (status NANORQ_SYM_IGN) */
#define NANORQ_TAG_NONE 0xFFFFFFFFu

/* ---- decoder: n symbols, symbol k = T bytes at data + k*pitch with tag tags[k]
 * (nanorq_tag; any mix of blocks and ESIs, any order; NANORQ_TAG_NONE = no symbol in this row).  Replaces n calls of
 * nanorq_decoder_add_symbol (lib/nanorq.c:478-509) and classifies every symbol the same way;
 * status[k] (optional) receives NANORQ_SYM_ADDED / _IGN / _DUP / _ERR.  Returns the number of
 * symbols added, or -1 if any symbol was rejected with NANORQ_SYM_ERR.
 *
 * Where the decoded bytes appear: source symbols are written to `io` when they arrive, like the
 * per-symbol call -- unless `io` is an ioctx_from_pinned_mem, in which case a block is written
 * as a whole (received and recovered symbols, one DMA from the device) by the call that
 * completes it: nanorq_repair_block, or this call when it delivers a block's last missing
 * source symbol.  `data` must stay valid until that call has returned or nanorq_free. */
int nanorq_decoder_add_symbols(nanorq *rq, const uint32_t *tags, const void *data, size_t pitch, size_t n,
                               int *status, struct ioctx *io);

/* ---- decoder: repair several blocks of the object with ONE solve launch per device.
 * Replaces n calls of nanorq_repair_block (lib/nanorq.c:591-631): every block's constraint matrix is
 * analysed on the host, the solves of all blocks on one device run as a single kernel launch
 * (gridDim.y = blocks), then the recovered symbols are written block by block.  ok[k] (optional)
 * receives what nanorq_repair_block would have returned for sbns[k]; returns how many are true. */
size_t nanorq_repair_blocks(nanorq *rq, struct ioctx *io, const uint8_t *sbns, size_t n, bool *ok);

/* ---- devices: source blocks of one object are independent, so block sbn is solved on device
 * sbn mod n_devices (the reference's per-block state: lib/nanorq.c:57,130-146).
 * n = 0 selects every visible CUDA device; the default is 1 (the process-wide device of
 * rqb_set_device).  Affects blocks created afterwards.  Returns the number in effect. */
int nanorq_set_devices(nanorq *rq, int n);

#ifdef __cplusplus
}
#endif
#endif
