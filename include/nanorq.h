/* nanorq.h -- encoder/decoder API of the B200 build.
 *
 * Same functions, argument meaning and return conventions as the reference's
 * include/nanorq.h:16-83 (implementation there: lib/nanorq.c:206-631), so
 * encode.c / decode.c / benchmark.c style programs compile against this header
 * unchanged.  The intermediate-symbol solve and every symbol combination run on
 * the GPU (see rqb200.h); there is no CPU fallback: without a CUDA device
 * nanorq_generate_symbols / nanorq_repair_block fail (return false).
 */
#ifndef NANORQ_H
#define NANORQ_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "io.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NANORQ_SYM_DUP 2
#define NANORQ_SYM_IGN 1
#define NANORQ_SYM_ADDED 0
#define NANORQ_SYM_ERR -1
#define NANORQ_MAX_TRANSFER 946270874880ULL /* ~881 GB */

typedef struct nanorq nanorq;

/* lib/nanorq.c:294 / :241 -- K xor Z non-zero selects the partitioning */
nanorq *nanorq_encoder_new(size_t len, uint16_t T, uint8_t Al);
nanorq *nanorq_encoder_new_ex(size_t len, uint16_t T, uint16_t K, uint16_t Z, uint8_t Al);
/* :206 load block sbn through io and compute its intermediate symbols */
bool nanorq_generate_symbols(nanorq *rq, uint8_t sbn, struct ioctx *io);
void nanorq_free(nanorq *rq); /* :298 */
uint64_t nanorq_oti_common(nanorq *rq);           /* :309  F<<24 | (T-1)            */
uint32_t nanorq_oti_scheme_specific(nanorq *rq);  /* :317  (Z-1)<<24 | (N-1)<<8 | Al */
size_t nanorq_transfer_length(nanorq *rq);        /* :332 */
size_t nanorq_symbol_size(nanorq *rq);            /* :334 */
size_t nanorq_blocks(nanorq *rq);                 /* :389 */
size_t nanorq_block_symbols(nanorq *rq, uint8_t sbn); /* :379 */
uint32_t nanorq_tag(uint8_t sbn, uint32_t esi);   /* :326 */
size_t nanorq_max_blocks(nanorq *rq);             /* :387 */
bool nanorq_precalculate(nanorq *rq);             /* :393 */
/* :403 returns T on success, 0 on failure */
size_t nanorq_encode(nanorq *rq, void *data, uint32_t esi, uint8_t sbn, struct ioctx *io);
void nanorq_encoder_cleanup(nanorq *rq, uint8_t sbn); /* :437 */
void nanorq_encoder_reset(nanorq *rq, uint8_t sbn);   /* :453 */
nanorq *nanorq_decoder_new(uint64_t common, uint32_t specific); /* :336 */
/* :471 widens the ESI range a decoder accepts (default 2 K'); call it before the first symbol of a block.
 * It does not widen the number of symbols a block holds: at most 2 K' + 1024 distinct symbols are kept per
 * block (a block is decodable long before that), further ones are refused with NANORQ_SYM_ERR */
bool nanorq_set_max_esi(nanorq *rq, uint32_t max_esi);
/* :478 returns NANORQ_SYM_* */
int nanorq_decoder_add_symbol(nanorq *rq, void *data, uint32_t tag, struct ioctx *io);
size_t nanorq_num_missing(nanorq *rq, uint8_t sbn); /* :511 */
size_t nanorq_num_repair(nanorq *rq, uint8_t sbn);  /* :519 */
/* :591 false => need more symbols; may be called again */
bool nanorq_repair_block(nanorq *rq, struct ioctx *io, uint8_t sbn);

#ifdef __cplusplus
}
#endif
#endif
