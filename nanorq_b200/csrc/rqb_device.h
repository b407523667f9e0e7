/* rqb_device.h -- the thin extern "C" shim between the C host code and the
 * sm_100a kernels in rqb_device.cu.  Plain pointers and sizes only.
 * Every function returns 0 on success or a cudaError_t value; the text of the
 * last failure is available from rqb_dev_last_error(). */
#ifndef RQB_DEVICE_H
#define RQB_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#include "rqb_rfc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one row operation, binary-compatible with the reference's sched_op
 * (include/sched.h:6-10): beta>=1: D[i] ^= beta*D[j] ; beta==0: D[i] *= (uint8_t)j */
typedef struct {
  uint8_t beta;
  uint32_t i;
  uint32_t j;
} rqb_rowop;

/* arguments of one source block for the column-sliced solve kernel
 * (all pointers are DEVICE pointers) */
typedef struct {
  uint8_t *base;        /* the block's arena: [IN | SYM | C | WS] rows (rqb_program.h)        */
  const uint8_t *pages; /* program pages                                                     */
  uint32_t pitch;       /* bytes between rows (all spaces), multiple of 64                   */
  uint32_t n_pages;
  uint32_t width;       /* bytes per row to process (multiple of 16)                         */
  uint32_t pad;
} rqb_solve_args;

int rqb_dev_count(void);
int rqb_dev_set(int dev);
int rqb_dev_get(void);
int rqb_dev_sm_count(void);
const char *rqb_dev_last_error(void);

int rqb_dev_malloc(void **p, size_t bytes);
int rqb_dev_free(void *p);
int rqb_host_malloc(void **p, size_t bytes); /* pinned */
int rqb_host_free(void *p);
int rqb_stream_create(void **s);
int rqb_stream_destroy(void *s);
int rqb_stream_sync(void *s);
/* wait for the stream through a flag word in pinned host memory (see rqb_device.cu) */
int rqb_stream_wait_flag(void *s, uint32_t *flag, uint32_t *seq);
int rqb_dev_sync(void);
int rqb_copy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int rqb_copy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int rqb_copy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int rqb_copy2d_h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                   size_t rows, void *stream);
int rqb_copy2d_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                   size_t rows, void *stream);
int rqb_dev_memset(void *p, int v, size_t bytes, void *stream);
int rqb_event_create(void **e);
int rqb_event_create_sync(void **e); /* no timing: ordering between streams only */
int rqb_stream_wait_event(void *stream, void *event);
int rqb_dev_mem_info(size_t *free_bytes, size_t *total_bytes);
/* page-lock caller memory so that copies to and from it are plain DMA */
int rqb_host_register(void *p, size_t bytes);
int rqb_host_unregister(void *p);
int rqb_event_destroy(void *e);
int rqb_event_record(void *e, void *stream);
int rqb_event_sync(void *e);
int rqb_event_elapsed_ms(void *start, void *stop, float *ms);

/* launches ONE kernel over nblocks source blocks (gridDim.y); args_dev is a
 * device array of rqb_solve_args; max_width is the maximum over the batch */
int rqb_launch_solve(const rqb_solve_args *args_dev, int nblocks, uint32_t max_width, void *stream);
/* the shared-memory flavour of the program (rqb_program.h): all blocks of the launch use
 * slice_bytes-wide slots (16, 32 or 64), at most max_slots of them */
int rqb_launch_solve_smem(const rqb_solve_args *args_dev, int nblocks, uint32_t max_width, uint32_t slice_bytes,
                          uint32_t max_slots, void *stream);
int rqb_smem_slot_budget_bytes(void);
/* column slice (bytes per CTA: 64, 128 or 256) rqb_launch_solve picks for such a launch */
int rqb_solve_slice_bytes(int nblocks, uint32_t max_width);

/* LT combine (decode_row, lib/nanorq.c:184-204): out[k] = XOR of the
 * intermediate symbols selected by Tuple[K', isi[k]]; tuples are generated on
 * the device.  isi_dev: device array. */
int rqb_launch_lt(const rqb_params *P, const uint8_t *c, uint32_t c_pitch, const uint32_t *isi_dev,
                  uint32_t n, uint8_t *out, uint32_t out_pitch, uint32_t width, void *stream);

/* batched row operations out of HBM (oaxpy/oaddrow/oscal, oblas_avx.c:43-114).
 * The ops of one call must be independent of each other (no op reads or writes a
 * row another op writes).  ops_dev: device array. */
int rqb_launch_rowops(uint8_t *D, size_t pitch, uint32_t width, const rqb_rowop *ops_dev,
                      uint32_t n, void *stream);
/* out-of-place row gather (precode_matrix_permute, lib/precode.c:3-13):
 * dst[k] = src[map[k]] */
int rqb_launch_gather_rows(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch,
                           const uint32_t *map_dev, uint32_t n, uint32_t width, void *stream);

/* row copies inside one arena: row pairs[2k+1] = row pairs[2k], k < n (pairs_dev: device array) */
int rqb_launch_copy_rows(uint8_t *base, size_t pitch, const uint32_t *pairs_dev, uint32_t n, uint32_t width,
                         void *stream);

/* dst row k = src row k for n rows of `width` bytes with different pitches, all on the device */
int rqb_launch_repitch(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, uint32_t width, uint32_t n,
                       void *stream);

/* the u x u Schur elimination of one block on the device (rqb_usolve_kernel): buf_dev = header +
 * matrices, laid out as rqb_solver.c's usolve hook packs them */
size_t rqb_usolve_buffer_bytes(int nb, int U, int uw, int nbw, int H, int sh_stride);
size_t rqb_usolve_header_bytes(void);
int rqb_launch_usolve(uint8_t *buf_dev, int nb, int U, int uw, int nbw, void *stream);

/* kernels launched by this process so far (bench.py's gpu_launches) */
unsigned long long rqb_dev_launch_count(void);
/* bytes moved by rqb_copy* so far */
void rqb_dev_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h);
/* process-wide default device for new contexts (CUDA's current device is per thread) */
void rqb_dev_set_default(int dev);
int rqb_dev_default(void);

#ifdef __cplusplus
}
#endif
#endif
