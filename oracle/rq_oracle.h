/* rq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99, scalar) of the reference hot path: RFC 6330
 * parameters/tuples, the precode matrix, the inactivation-style schedule
 * construction, the schedule replay on the symbol matrix D, and the LT
 * row-combine.  It follows the reference's algorithm step for step (same
 * peeling order, same pivot order, same op list) so that op counts and op
 * sequences can be compared 1:1 with the compiled reference (oracle/_ref).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this.  The product (nanorq_b200/) never does.
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks this file against the
 * unmodified reference compiled from /root/reference (oracle/_ref) and
 * against the committed KAT fixtures in tests/golden/.
 */
#ifndef RQ_ORACLE_H
#define RQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int Kprime, S, H, W, L, P, P1, U, B, J;
} orc_params;

typedef struct {
  uint32_t d, a, b, d1, a1, b1;
} orc_tuple;

/* one recorded row operation; same meaning as the reference's sched_op
 * (include/sched.h:6-10): beta>=1 => D[i] ^= beta*D[j]; beta==0 => D[i] *= j */
typedef struct {
  uint8_t beta;
  uint32_t i, j;
} orc_op;

typedef struct {
  int rows, cols;
  int *c, *ci, *d, *di;
  orc_op *ops;
  size_t nops, cap;
  int i, u;
  long marks[2];
  int promotions; /* wrkmat GF2->GF256 row promotions that happened */
} orc_sched;

/* lib/params.c:21-45 */
int orc_params_init(int K, orc_params *out);
/* lib/rand.c:183-190 */
uint32_t orc_rand(uint32_t y, uint32_t i, uint32_t m);
/* lib/tuple.c:21-43 */
orc_tuple orc_tuple_gen(const orc_params *P, uint32_t X);
/* lib/params.c:47-65 ; returns number of indices written (<= 33) */
int orc_lt_indices(const orc_params *P, uint32_t X, uint32_t *out);

/* GF(256), polynomial 0x11D (deps/oblas/tablegen.c:10,31-52) */
uint8_t orc_gf_mul(uint8_t a, uint8_t b);
uint8_t orc_gf_inv(uint8_t a);
uint8_t orc_gf_exp(int e);
/* deps/oblas/oblas_classic.c semantics of oaxpy / oscal on one row of n bytes */
void orc_row_axpy(uint8_t *dst, const uint8_t *src, size_t n, uint8_t u);
void orc_row_scal(uint8_t *dst, size_t n, uint8_t u);
/* batched form used by the row-op parity tests: ops applied in order on a
 * row-major matrix with the given pitch */
void orc_apply_ops(uint8_t *D, size_t pitch, size_t T, const orc_op *ops,
                   size_t nops);

/* precode_matrix_gen + patch_precode_matrix + precode_matrix_invert
 * (lib/precode.c:90-97,347-377 ; lib/nanorq.c:527-547).
 * isi[k], k in [0, Kprime+overhead): the internal symbol id whose LT row sits
 * in matrix row S+H+k.  Returns NULL when rank < L (reference returns NULL too).
 * *status: 0 ok, 1 singular, 2 the reference would abort() in wrkmat. */
orc_sched *orc_invert(const orc_params *P, int overhead, const uint32_t *isi,
                      int *status);
void orc_sched_free(orc_sched *S);
/* number of row ops precode_matrix_apply_sched performs (lib/precode.c:23-32) */
size_t orc_applied_ops(const orc_sched *S, size_t *n_axpy, size_t *n_scal);
/* precode_matrix_intermediate (lib/precode.c:379-389): replay + two permutes */
void orc_intermediate(const orc_sched *S, uint8_t *D, size_t pitch, size_t T);
/* decode_row (lib/nanorq.c:184-204) */
void orc_lt_row(const orc_params *P, const uint8_t *C, size_t pitch,
                uint32_t isi, uint8_t *out, size_t T);

uint64_t orc_fnv1a64(const uint8_t *p, size_t n);

/* Whole-block helpers mirroring nanorq_generate_symbols / nanorq_repair_block
 * (lib/nanorq.c:206-232, 591-631).
 *  - encode: src = K*T bytes (row-major), C_out = L rows of pitch bytes.
 *  - decode: esis/syms in ARRIVAL order (n symbols, T bytes each); recovered
 *    K*T source bytes to out; C_out optional (L rows).  Returns 0 ok,
 *    1 = need more symbols / singular, 2 = reference would abort. */
int orc_encode_block(int K, size_t T, const uint8_t *src, uint8_t *C_out,
                     size_t pitch, size_t *nops, size_t *n_applied);
int orc_decode_block(int K, size_t T, const uint32_t *esis,
                     const uint8_t *syms, size_t n, uint8_t *out,
                     uint8_t *C_out, size_t pitch, size_t *nops,
                     size_t *n_applied);

#ifdef __cplusplus
}
#endif
#endif
