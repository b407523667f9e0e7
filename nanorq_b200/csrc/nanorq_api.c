/* nanorq_api.c -- the nanorq.h encoder/decoder object on top of the GPU block
 * solver (rqb200.h).  Host glue only: OTI packing, source-block partitioning,
 * symbol bookkeeping and ioctx traffic follow the reference's lib/nanorq.c
 * (cited per function); all symbol arithmetic happens on the device.
 */
#include "nanorq.h"
#include "nanorq_batch.h"

#include <stdlib.h>
#include <string.h>

#include "rqb200.h"
#include "rqb_hostcopy.h"
#include "rqb_prof.h"

#ifndef UPLOAD_CHUNK
#define UPLOAD_CHUNK 512u /* rows per host->device copy queued while the block is still being read */
#endif
#ifndef WINDOW_DIV
#define WINDOW_DIV 8u /* repair symbols produced per device window: K / WINDOW_DIV (32..8192) */
#endif

#define Z_MAX 256
#define K_MAX 56403

struct part {
  size_t IL, IS, JL, JS; /* long size, short size, #long, #short */
};

struct block {
  uint16_t K;
  bool loaded, inverted;
  rqb_solver *sv;
  size_t pitch;
  /* decoder bookkeeping (lib/nanorq.c:40-47: repair_bin, repair_mask) */
  uint32_t *mask;
  size_t mask_words;
  size_t gaps;
  uint32_t *src_row;           /* [K] input row holding source symbol esi (valid where the mask is set) */
  uint32_t *rep_esi, *rep_row; /* repair symbols in arrival order: ESI and input row */
  size_t nrep, rep_cap;
  uint32_t in_cap;
  uint32_t landed;               /* input rows handed out so far (arrival order) */
  uint32_t staged_lo, staged_hi; /* staging rows written by per-symbol calls, not yet queued for upload */
  bool out_decided, written;     /* output mode chosen / deferred block image already written */
  uint8_t *out_mem;              /* deferred output: the block's bytes inside a page-locked memory ioctx */
  size_t out_bytes;
  /* encoder: window of repair symbols already produced on the device */
  uint32_t win_first, win_n, win_cap;
  bool win_pending; /* the window's copy to the host is queued but not waited for */
  bool win_on_host; /* the pinned mirror holds the current window */
  bool deferred;    /* generate_symbols was asked for, the solve is queued when a repair symbol is needed */
  uint32_t loaded_rows; /* staging rows already queued for upload */
  const uint8_t *src_mem; /* encoder loaded by DMA from a page-locked memory ioctx: the block's bytes there */
  size_t src_bytes;
  /* small blocks get their device context only when a solve is needed (need_ctx): until then their
   * symbols sit in plain host rows -- a transfer without loss never touches the GPU */
  uint8_t *lazy_stage;
  uint32_t max_out;
  int dev;
};

struct nanorq {
  size_t F, T, Al, Z, N, Kt;
  struct part src_part, sub_part;
  rqb_block_params P;
  uint32_t max_esi;
  int n_dev; /* devices the blocks are spread over (nanorq_set_devices); 0/1 = the default device */
  struct block *blocks[Z_MAX];
};

static size_t div_ceil(size_t a, size_t b) { return a / b + (a % b ? 1 : 0); }

static struct part make_part(size_t I, size_t J) { /* lib/nanorq.c:83-95 */
  struct part p = {0, 0, 0, 0};
  if (J == 0) return p;
  p.IL = div_ceil(I, J);
  p.IS = I / J;
  p.JL = I - p.IS * J;
  p.JS = J - p.JL;
  if (p.JL == 0) p.IL = 0;
  return p;
}

size_t nanorq_block_symbols(nanorq *rq, uint8_t sbn) { /* :379-385 */
  if (sbn < rq->src_part.JL) return rq->src_part.IL;
  if (sbn - rq->src_part.JL < rq->src_part.JS) return rq->src_part.IS;
  return 0;
}

size_t nanorq_max_blocks(nanorq *rq) {
  (void)rq;
  return Z_MAX;
}
size_t nanorq_blocks(nanorq *rq) { return rq->src_part.JL + rq->src_part.JS; }
size_t nanorq_transfer_length(nanorq *rq) { return rq->F; }
size_t nanorq_symbol_size(nanorq *rq) { return rq->T; }

uint64_t nanorq_oti_common(nanorq *rq) { /* :309-315 */
  return ((uint64_t)rq->F << 24) | ((rq->T - 1) & 0xffff);
}

uint32_t nanorq_oti_scheme_specific(nanorq *rq) { /* :317-324 */
  return (uint32_t)((rq->Z - 1) << 24) | (uint32_t)((rq->N - 1) << 8) | (uint32_t)rq->Al;
}

uint32_t nanorq_tag(uint8_t sbn, uint32_t esi) { return ((uint32_t)sbn << 24) | (esi & 0x00ffffff); }

nanorq *nanorq_encoder_new_ex(size_t len, uint16_t T16, uint16_t K, uint16_t Z16, uint8_t Al) {
  /* lib/nanorq.c:241-292 and gen_scheme_specific :60-81 */
  static const uint8_t aligns[] = {1, 2, 4, 8};
  if (len == 0 || len > NANORQ_MAX_TRANSFER) return NULL;
  uint8_t al = 1;
  for (int a = 3; a >= 0; a--)
    if (Al >= aligns[a]) {
      al = aligns[a];
      break;
    }
  size_t T = T16, Z = Z16;
  if (T < al)
    T = al;
  else
    T -= T % al;
  while (div_ceil(len, T) > (size_t)Z_MAX * K_MAX) {
    if (al == 1) return NULL; /* the reference would spin forever here */
    T *= al;
    if (T > 65535) return NULL;
  }
  size_t Kt = div_ceil(len, T), Kn = K;
  if (K == 0) {
    Kn = Kt;
    if (Z == 0) {
      Z = 16;
      while (div_ceil(Kt, Z) > K_MAX) Z++;
    }
    Kn = div_ceil(Kt, Z);
  }
  if (Kn == 0) return NULL;
  size_t Zf = div_ceil(Kt, Kn);
  if (Zf == 0 || Zf > Z_MAX || div_ceil(Kt, Zf) > K_MAX) return NULL;
  nanorq *rq = calloc(1, sizeof(*rq));
  if (!rq) return NULL;
  rq->F = len;
  rq->T = T;
  rq->Al = al;
  rq->Kt = Kt;
  rq->Z = Zf;
  rq->N = 1; /* sub-block interleaving is disabled in the reference as well (:78) */
  rq->src_part = make_part(Kt, Zf);
  rq->sub_part = make_part(T / al, 1);
  if (rqb_block_params_init((int)nanorq_block_symbols(rq, 0), &rq->P)) {
    free(rq);
    return NULL;
  }
  return rq;
}

nanorq *nanorq_encoder_new(size_t len, uint16_t T, uint8_t Al) { return nanorq_encoder_new_ex(len, T, 0, 0, Al); }

nanorq *nanorq_decoder_new(uint64_t common, uint32_t scheme) { /* :336-376 */
  uint64_t F = common >> 24;
  size_t T = (size_t)(common & 0xffff) + 1;
  if (F == 0 || F > NANORQ_MAX_TRANSFER) return NULL;
  size_t Z = ((scheme >> 24) & 0xff) + 1, N = ((scheme >> 8) & 0xffff) + 1, Al = scheme & 0xff;
  if (Al == 0 || T < Al || T % Al != 0) return NULL;
  /* Sub-block interleaving: the reference's encoder never produces N > 1 (gen_scheme_specific,
   * lib/nanorq.c:78 "disable interleaving"); its decoder would walk the sub-blocks of such an OTI
   * (:97-173), this one lays symbols out contiguously, so a foreign or corrupted N is refused
   * rather than decoded to the wrong offsets.  With N = 1 the value of Al only has to divide T. */
  if (N != 1) return NULL;
  size_t Kt = div_ceil(F, T);
  if (div_ceil(Kt, Z) > K_MAX) return NULL;
  nanorq *rq = calloc(1, sizeof(*rq));
  if (!rq) return NULL;
  rq->F = F;
  rq->T = T;
  rq->Al = Al;
  rq->Z = Z;
  rq->N = N;
  rq->Kt = Kt;
  rq->src_part = make_part(Kt, Z);
  rq->sub_part = make_part(T / Al, N);
  if (nanorq_block_symbols(rq, 0) == 0 || rqb_block_params_init((int)nanorq_block_symbols(rq, 0), &rq->P)) {
    free(rq);
    return NULL;
  }
  rq->max_esi = 2u * (uint32_t)rq->P.Kprime;
  return rq;
}

bool nanorq_set_max_esi(nanorq *rq, uint32_t max_esi) { /* :471-476 */
  if (!rq || max_esi >= (1u << 24) || max_esi < (uint32_t)rq->P.Kprime) return false;
  rq->max_esi = max_esi;
  return true;
}

/* byte offset of symbol esi of block sbn in the object (N = 1: one contiguous
 * span, get_source_block / get_symbol_offset :97-128) */
static size_t symbol_offset(nanorq *rq, uint8_t sbn, uint32_t esi) {
  size_t first;
  if (sbn < rq->src_part.JL)
    first = (size_t)sbn * rq->src_part.IL;
  else
    first = rq->src_part.IL * rq->src_part.JL + ((size_t)sbn - rq->src_part.JL) * rq->src_part.IS;
  return (first + esi) * rq->T;
}

/* transfer_esi :148-173 for N = 1 */
static size_t transfer_symbol(nanorq *rq, uint8_t sbn, uint32_t esi, uint8_t *ptr, struct ioctx *io, int out) {
  size_t off = symbol_offset(rq, sbn, esi), n = rq->T;
  if (off >= rq->F) return 0;
  if (!io->seek(io, off)) return 0;
  if (off + n >= rq->F) n = rq->F - off;
  return out ? io->write(io, ptr, n) : io->read(io, ptr, n);
}

static void block_free(struct block *b) {
  if (!b) return;
  rqb_solver_destroy(b->sv);
  free(b->lazy_stage);
  free(b->mask);
  free(b->rep_esi);
  free(b->rep_row);
  free(b->src_row);
  free(b);
}

/* first symbol of block sbn in the object, in symbols (get_source_block :97-112) */
static size_t block_first_symbol(nanorq *rq, uint8_t sbn) {
  if (sbn < rq->src_part.JL) return (size_t)sbn * rq->src_part.IL;
  return rq->src_part.IL * rq->src_part.JL + ((size_t)sbn - rq->src_part.JL) * rq->src_part.IS;
}

/* the block's span inside a memory ioctx this library made (ioctx_from_mem / _from_pinned_mem):
 * *bytes = payload bytes of the block present in the object (the object's last symbol may be
 * short); returns the address of the block's first byte or NULL */
int rqb_ioctx_mem_view(struct ioctx *io, uint8_t **base, size_t *len, int *pinned);
static uint8_t *block_span(nanorq *rq, uint8_t sbn, const struct block *b, struct ioctx *io, size_t *bytes,
                           int *pinned) {
  uint8_t *base = NULL;
  size_t len = 0;
  if (!rqb_ioctx_mem_view(io, &base, &len, pinned)) return NULL;
  const size_t off = block_first_symbol(rq, sbn) * rq->T;
  if (off >= rq->F || rq->F > len) return NULL;
  size_t n = (size_t)b->K * rq->T;
  if (off + n > rq->F) n = rq->F - off;
  *bytes = n;
  return base + off;
}

#define LAZY_MAX_SYMBOLS 256u /* blocks up to this size get a device context, and are solved, on demand */
#define LAZY_MAX_BYTES (256u << 10)

/* the host rows a block's symbols are collected in: the context's pinned staging area, or plain
 * memory while the block has no context yet */
static inline uint8_t *stage(const struct block *b) { return b->sv ? rqb_solver_staging(b->sv) : b->lazy_stage; }

/* the block needs the GPU now: take a device context and move what was collected so far into its
 * staging rows (nothing of it has been uploaded yet) */
static bool need_ctx(nanorq *rq, struct block *b) {
  if (b->sv) return true;
  if (rqb_solver_create_on(&b->sv, b->dev, (int)b->K, rq->P.Kprime, rq->T, b->in_cap, b->max_out) != 0) {
    b->sv = NULL;
    return false;
  }
  if (b->lazy_stage) {
    const uint32_t rows = b->mask ? b->landed : (b->loaded ? b->K : 0);
    memcpy(rqb_solver_staging(b->sv), b->lazy_stage, (size_t)rows * b->pitch);
    free(b->lazy_stage);
    b->lazy_stage = NULL;
    b->loaded_rows = 0;
    b->staged_lo = 0;
    b->staged_hi = b->mask ? b->landed : 0;
  }
  return true;
}

static struct block *get_block(nanorq *rq, uint8_t sbn) { /* get_block_encoder :130-146 */
  if (rq->blocks[sbn]) return rq->blocks[sbn];
  size_t K = nanorq_block_symbols(rq, sbn);
  if (K == 0) return NULL;
  struct block *b = calloc(1, sizeof(*b));
  if (!b) return NULL;
  b->K = (uint16_t)K;
  uint32_t max_in, max_out;
  if (rq->max_esi) { /* decoder: every symbol that may arrive lands in its own input row */
    uint32_t spare = rq->max_esi >= K ? rq->max_esi - (uint32_t)K + 1 : 1;
    /* nanorq_set_max_esi widens the ESI range, not the number of symbols a block can hold: a block
     * is decodable long before it has collected K' repair symbols beyond its own size */
    if (spare > (uint32_t)rq->P.Kprime + 1024u) spare = (uint32_t)rq->P.Kprime + 1024u;
    max_in = (uint32_t)rq->P.Kprime + spare;
    max_out = (uint32_t)K;
    b->mask_words = rq->max_esi / 32 + 2;
    b->mask = calloc(b->mask_words, sizeof(uint32_t));
    b->src_row = malloc(sizeof(uint32_t) * K);
    if (!b->mask || !b->src_row) {
      block_free(b);
      return NULL;
    }
    b->gaps = K;
  } else {
    max_in = (uint32_t)K;
    b->win_cap = (uint32_t)(K / WINDOW_DIV < 32 ? 32 : (K / WINDOW_DIV > 8192 ? 8192 : K / WINDOW_DIV));
    max_out = b->win_cap;
  }
  b->in_cap = max_in;
  b->max_out = max_out;
  /* independent source blocks shard over the devices of the box, block sbn on device sbn mod n
   * (SURVEY 8(e)); one device unless nanorq_set_devices asked for more */
  b->dev = rq->n_dev > 1 ? (int)(sbn % (unsigned)rq->n_dev) : -1;
  b->pitch = (rq->T + 63) / 64 * 64; /* = rqb_solver_pitch of the context this block gets */
  if (K <= LAZY_MAX_SYMBOLS && K * rq->T <= LAZY_MAX_BYTES) {
    if (rqb_device_count() <= 0) { /* no device: fail here like a block that needs its context at once */
      block_free(b);
      return NULL;
    }
    b->lazy_stage = calloc((size_t)max_in, b->pitch); /* zeroed: the pad bytes of a row travel with it */
    if (!b->lazy_stage) {
      block_free(b);
      return NULL;
    }
  } else if (!need_ctx(rq, b)) {
    block_free(b);
    return NULL;
  }
  rq->blocks[sbn] = b;
  return b;
}

int nanorq_set_devices(nanorq *rq, int n) {
  const int have = rqb_device_count();
  if (!rq || have <= 0) return 0;
  if (n <= 0 || n > have) n = have;
  rq->n_dev = n;
  return n;
}

/* ------------------------------------------------------------------ encoder */
static bool load_block(nanorq *rq, uint8_t sbn, struct block *b, struct ioctx *io) {
  b->loaded_rows = 0;
  int pinned = 0;
  size_t bytes = 0;
  uint8_t *span = block_span(rq, sbn, b, io, &bytes, &pinned);
  if (span && pinned && b->sv) { /* (a block without a context yet is small: its rows are copied by the CPU) */
    /* page-locked caller memory: the copy engine reads the block where it lies (no staging copy);
     * only a short last symbol goes through a zero-padded staging row */
    const uint32_t full = (uint32_t)(bytes / rq->T);
    if (rqb_solver_upload_rows(b->sv, 0, full, span, rq->T)) return false;
    if (full < b->K) {
      uint8_t *row = rqb_solver_staging(b->sv) + (size_t)full * b->pitch;
      memcpy(row, span + (size_t)full * rq->T, bytes - (size_t)full * rq->T);
      memset(row + (bytes - (size_t)full * rq->T), 0, rq->T - (bytes - (size_t)full * rq->T));
      for (uint32_t r = full + 1; r < b->K; r++) memset(rqb_solver_staging(b->sv) + (size_t)r * b->pitch, 0, rq->T);
      if (rqb_solver_upload(b->sv, full, b->K - full)) return false;
    }
    b->src_mem = span;
    b->src_bytes = bytes;
    b->loaded_rows = b->K;
    return true;
  }
  /* load_symbol_matrix :175-182: K reads of one symbol each into the (pinned) staging
   * rows; the rows start moving to the GPU while the rest is still being read */
  uint8_t *st = stage(b);
  b->src_mem = NULL;
  for (uint32_t esi = 0; esi < b->K; esi++) {
    uint8_t *row = st + (size_t)esi * b->pitch;
    size_t got = transfer_symbol(rq, sbn, esi, row, io, 0);
    if (got < rq->T) memset(row + got, 0, rq->T - got);
    if (b->sv && esi + 1 - b->loaded_rows == UPLOAD_CHUNK) {
      if (rqb_solver_upload(b->sv, b->loaded_rows, UPLOAD_CHUNK)) return false;
      b->loaded_rows = esi + 1;
    }
  }
  return true;
}

/* queue the upload of what is still in the staging rows, the solve and the first window of repair
 * symbols (the program is cached per K: cf. rq->S :219-221); nothing here waits for the device */
static bool launch_solve(nanorq *rq, struct block *b) {
  PF_T0;
  if (!need_ctx(rq, b)) return false;
  if (rqb_solver_upload(b->sv, b->loaded_rows, b->K - b->loaded_rows)) return false;
  b->loaded_rows = b->K;
  PF(RQB_PF_GEN_UPLOAD);
  if (rqb_solver_plan_encode(b->sv, 1, b->win_cap)) return false;
  PF(RQB_PF_GEN_PLAN);
  if (rqb_solver_run(b->sv)) return false;
  PF(RQB_PF_GEN_RUN);
  b->inverted = true;
  b->deferred = false;
  b->win_first = b->K;
  b->win_n = b->win_cap;
  b->win_pending = false;
  b->win_on_host = false; /* the window stays on the device until a per-symbol call asks for one of its symbols */
  return true;
}

/* Small blocks are solved on demand: a solve that nobody needs -- every transfer without loss -- costs a
 * launch and a stream wait, more than all the host work on a block of a few KB, and when repair
 * symbols are asked for after all a solve this small takes tens of microseconds.  Large blocks are
 * solved right away so that the solve overlaps the emission of the source symbols. */
bool nanorq_generate_symbols(nanorq *rq, uint8_t sbn, struct ioctx *io) { /* :206-232 */
  struct block *b = get_block(rq, sbn);
  if (!b) return false;
  if (b->inverted || b->deferred) return true;
  PF_T0;
  if (!b->loaded) b->loaded = load_block(rq, sbn, b, io);
  if (!b->loaded) return false;
  PF(RQB_PF_GEN_LOAD);
  if (!b->sv) { /* a small block: solved when a repair symbol is asked for */
    b->deferred = true;
    return true;
  }
  return launch_solve(rq, b);
}

/* the intermediate symbols are needed now */
static bool ensure_solved(nanorq *rq, uint8_t sbn, struct block *b, struct ioctx *io) {
  if (b->inverted) return true;
  if (!b->deferred && !nanorq_generate_symbols(rq, sbn, io)) return false;
  return b->inverted || launch_solve(rq, b);
}

bool nanorq_precalculate(nanorq *rq) { /* :393-401 */
  struct block *b = get_block(rq, 0);
  if (!b) return false;
  if (!b->sv) return true; /* small blocks: the program is built (and cached per K) with the first solve */
  return rqb_solver_plan_encode(b->sv, 1, b->win_cap) == 0;
}

/* source symbol esi as the block was loaded (zero-padded past the end of the object) */
static void copy_source_symbol(nanorq *rq, const struct block *b, uint32_t esi, uint8_t *dst) {
  if (!b->src_mem) {
    memcpy(dst, stage(b) + (size_t)esi * b->pitch, rq->T);
    return;
  }
  const size_t off = (size_t)esi * rq->T;
  const size_t have = off >= b->src_bytes ? 0 : (b->src_bytes - off < rq->T ? b->src_bytes - off : rq->T);
  memcpy(dst, b->src_mem + off, have);
  if (have < rq->T) memset(dst + have, 0, rq->T - have);
}

/* produce repair symbols esi .. esi+n-1 (n <= win_cap) on the device in one LT launch; they stay
 * in the emitted-symbol rows 0..n-1 */
static bool emit_window(nanorq *rq, struct block *b, uint32_t esi, uint32_t n) {
  const uint32_t pad = (uint32_t)rq->P.Kprime - b->K;
  uint32_t *isi = malloc(sizeof(uint32_t) * n);
  if (!isi) return false;
  for (uint32_t k = 0; k < n; k++) isi[k] = esi + k + pad; /* ISI = ESI + (K'-K) :429 */
  int rc = rqb_solver_emit(b->sv, isi, n);
  free(isi);
  if (rc) return false;
  b->win_first = esi;
  b->win_n = n;
  b->win_on_host = false;
  b->win_pending = false; /* the mirror copy queued by generate_symbols belongs to a window that is gone */
  return true;
}

size_t nanorq_encode(nanorq *rq, void *data, uint32_t esi, uint8_t sbn, struct ioctx *io) { /* :403-435 */
  struct block *b = get_block(rq, sbn);
  if (!b) return 0;
  if (esi < b->K) {
    /* source symbol: the bytes the block was loaded with (the reference re-derives
     * them from the intermediate symbols once inverted; same bytes) */
    if (!b->loaded) b->loaded = load_block(rq, sbn, b, io);
    if (!b->loaded) return 0;
    PF_T0;
    copy_source_symbol(rq, b, esi, data);
    PF(RQB_PF_EMIT_SRC);
    return rq->T;
  }
  if (esi > ((1u << 24) - 1)) return 0;
  if (!ensure_solved(rq, sbn, b, io)) return 0;
  PF_T0;
  if (b->win_pending) { /* the solve and the copy of the first window were queued by generate_symbols */
    if (rqb_solver_sync(b->sv)) return 0;
    b->win_pending = false;
    b->win_on_host = true;
    PF(RQB_PF_GEN_SYNC);
  }
  if (!(b->win_n && esi >= b->win_first && esi < b->win_first + b->win_n)) {
    uint32_t n = b->win_cap;
    if ((uint64_t)esi + n > (1u << 24)) n = (1u << 24) - esi;
    if (!emit_window(rq, b, esi, n)) return 0;
  }
  if (!b->win_on_host) { /* bring the window to the pinned mirror once; per-symbol calls are views of it */
    if (rqb_solver_fetch_syms(b->sv, 0, b->win_n, NULL, 0)) return 0;
    b->win_on_host = true;
  }
  memcpy(data, rqb_solver_sym_mirror(b->sv) + (size_t)(esi - b->win_first) * b->pitch, rq->T);
  PF(RQB_PF_EMIT_WINDOW);
  return rq->T;
}

size_t nanorq_encode_range(nanorq *rq, uint8_t sbn, uint32_t esi0, uint32_t n, void *dst, size_t pitch,
                           struct ioctx *io) {
  struct block *b = rq ? get_block(rq, sbn) : NULL;
  if (!b || b->mask || !dst || pitch < rq->T || n == 0 || (uint64_t)esi0 + n > (1u << 24)) return 0;
  PF_T0;
  if (esi0 + n <= b->K && !b->inverted) { /* source symbols only and nothing solved yet: they are the loaded bytes */
    if (!b->loaded) b->loaded = load_block(rq, sbn, b, io);
    if (!b->loaded) return 0;
    if (b->K <= LAZY_MAX_SYMBOLS) {
      for (uint32_t k = 0; k < n; k++) copy_source_symbol(rq, b, esi0 + k, (uint8_t *)dst + (size_t)k * pitch);
      return n;
    }
  }
  if (!ensure_solved(rq, sbn, b, io)) return 0;
  PF(RQB_PF_RANGE_GEN);
  uint8_t *out = dst;
  uint32_t esi = esi0, left = n;
  /* source symbols: straight from the device's input rows into the caller's rows (DMA when the
   * destination is page-locked) -- the host never touches the bytes */
  if (esi < b->K) {
    const uint32_t m = b->K - esi < left ? b->K - esi : left;
    if (rqb_solver_fetch_rows(b->sv, 0, esi, m, out, pitch, 0)) return 0;
    out += (size_t)m * pitch;
    esi += m;
    left -= m;
  }
  /* repair symbols: whatever part of the range the current device window holds is copied from
   * there, the rest is produced window by window (one LT launch each) */
  while (left) {
    if (!(b->win_n && esi >= b->win_first && esi < b->win_first + b->win_n)) {
      uint32_t w = left < b->win_cap ? left : b->win_cap;
      if (!emit_window(rq, b, esi, w)) return 0; /* queued behind whatever still reads the old window's rows */
    }
    const uint32_t avail = b->win_first + b->win_n - esi, m = avail < left ? avail : left;
    if (rqb_solver_fetch_rows(b->sv, 1, esi - b->win_first, m, out, pitch, 0)) return 0;
    out += (size_t)m * pitch;
    esi += m;
    left -= m;
  }
  PF(RQB_PF_RANGE_QUEUE);
  if (rqb_solver_sync(b->sv)) return 0;
  PF(RQB_PF_RANGE_WAIT);
  return n;
}

void nanorq_encoder_cleanup(nanorq *rq, uint8_t sbn) { /* :437-451 */
  if (!rq->blocks[sbn]) return;
  block_free(rq->blocks[sbn]);
  rq->blocks[sbn] = NULL;
}

void nanorq_encoder_reset(nanorq *rq, uint8_t sbn) { /* :453-469 */
  struct block *b = rq->blocks[sbn];
  if (!b) return;
  if (b->sv) rqb_solver_sync(b->sv); /* the staging rows are about to be rewritten */
  b->loaded = b->inverted = b->win_pending = b->win_on_host = b->deferred = false;
  b->win_n = 0;
  b->loaded_rows = 0;
  b->src_mem = NULL;
  b->nrep = 0;
  b->landed = b->staged_lo = b->staged_hi = 0;
  b->written = b->out_decided = false; /* the next round may come with another output ioctx */
  b->out_mem = NULL;
  b->out_bytes = 0;
  if (b->mask) {
    memset(b->mask, 0, b->mask_words * sizeof(uint32_t));
    b->gaps = b->K;
  }
}

void nanorq_free(nanorq *rq) { /* :298-307 (NULL-safe here) */
  if (!rq) return;
  PF_T0;
  rqb_copy_fence();
  for (int sbn = 0; sbn < Z_MAX; sbn++) nanorq_encoder_cleanup(rq, (uint8_t)sbn);
  free(rq);
  PF(RQB_PF_FREE);
}

/* ------------------------------------------------------------------ decoder
 * Every symbol a block accepts lands in its own input row, in arrival order (src_row[esi] /
 * rep_row[k] remember where): a whole batch of symbols is then ONE copy into consecutive rows,
 * straight from the caller's memory when it comes through nanorq_decoder_add_symbols, through
 * the pinned staging rows when it comes one symbol at a time. */
static inline bool mask_get(const struct block *b, uint32_t id) { return (b->mask[id / 32] >> (id % 32)) & 1; }
static inline void mask_set(struct block *b, uint32_t id) { b->mask[id / 32] |= 1u << (id % 32); }

/* staging rows [staged_lo, staged_hi) hold symbols not yet queued for upload */
static bool flush_staged(struct block *b) {
  if (!b->sv) return true; /* no context yet: the rows wait in plain memory (need_ctx moves them) */
  if (b->staged_hi > b->staged_lo && rqb_solver_upload(b->sv, b->staged_lo, b->staged_hi - b->staged_lo)) return false;
  b->staged_lo = b->staged_hi = b->landed;
  return true;
}

static bool grow_rep(struct block *b) {
  if (b->nrep < b->rep_cap) return true;
  size_t cap = b->rep_cap ? b->rep_cap * 2 : 256;
  uint32_t *e = realloc(b->rep_esi, cap * sizeof(uint32_t));
  if (e) b->rep_esi = e;
  uint32_t *r = realloc(b->rep_row, cap * sizeof(uint32_t));
  if (r) b->rep_row = r;
  if (!e || !r) return false; /* what was collected so far stays valid */
  b->rep_cap = cap;
  return true;
}

/* bookkeeping of one accepted symbol that sits in input row `row` (:478-509 without the copies) */
static int note_symbol(struct block *b, uint32_t esi, uint32_t row) {
  if (esi < b->K) {
    b->src_row[esi] = row;
    b->gaps--;
  } else {
    if (!grow_rep(b)) return NANORQ_SYM_ERR;
    b->rep_esi[b->nrep] = esi;
    b->rep_row[b->nrep++] = row;
  }
  mask_set(b, esi);
  return NANORQ_SYM_ADDED;
}

/* classification shared by the per-symbol and the batch call; SYM_ADDED = "would be added" */
static int classify(nanorq *rq, const struct block *b, uint32_t esi) {
  if (!b || !b->mask || esi > rq->max_esi || esi / 32 >= b->mask_words) return NANORQ_SYM_ERR;
  if (b->gaps == 0) return NANORQ_SYM_IGN;
  if (mask_get(b, esi)) return NANORQ_SYM_DUP;
  if (b->landed >= b->in_cap) return NANORQ_SYM_ERR;
  return NANORQ_SYM_ADDED;
}

static bool complete_block(nanorq *rq, struct ioctx *io, uint8_t sbn, struct block *b);

/* is the output a page-locked memory ioctx, so that a block is handed back as one DMA? */
static bool deferred_output(nanorq *rq, uint8_t sbn, struct block *b, struct ioctx *io) {
  if (b->out_decided) return b->out_mem != NULL;
  int pinned = 0;
  size_t bytes = 0;
  uint8_t *span = block_span(rq, sbn, b, io, &bytes, &pinned);
  b->out_decided = true;
  /* a block of a few KB is written by the CPU as its symbols arrive: handing it back through the
   * device would cost a kernel, a copy and a stream wait for less than a page of data */
  if (span && pinned && bytes > LAZY_MAX_BYTES) {
    b->out_mem = span;
    b->out_bytes = bytes;
  }
  return b->out_mem != NULL;
}

int nanorq_decoder_add_symbol(nanorq *rq, void *data, uint32_t tag, struct ioctx *io) { /* :478-509 */
  uint8_t sbn = (tag >> 24) & 0xff;
  uint32_t esi = tag & 0x00ffffff;
  PF_T0;
  struct block *b = get_block(rq, sbn);
  PF(RQB_PF_ADD_CREATE);
  int st = classify(rq, b, esi);
  if (st != NANORQ_SYM_ADDED) return st;
  const uint32_t row = b->landed;
  if (b->sv)
    rqb_copy_stream(rqb_solver_staging(b->sv) + (size_t)row * b->pitch, data, rq->T);
  else
    memcpy(b->lazy_stage + (size_t)row * b->pitch, data, rq->T);
  PF(RQB_PF_ADD_COPY);
  if (b->staged_hi != row) { /* rows of a batch call lie in between: start a new pending range */
    if (!flush_staged(b)) return NANORQ_SYM_ERR;
  }
  st = note_symbol(b, esi, row);
  if (st != NANORQ_SYM_ADDED) return st;
  b->landed = b->staged_hi = row + 1;
  if (esi < b->K) {
    if (!deferred_output(rq, sbn, b, io)) {
      transfer_symbol(rq, sbn, esi, data, io, 1); /* source symbols go straight to the output */
      PF(RQB_PF_ADD_WRITE);
      if (b->gaps == 0) rqb_copy_fence(); /* the block's output is complete: publish the streamed rows */
    } else if (b->gaps == 0 && !complete_block(rq, io, sbn, b)) {
      return NANORQ_SYM_ERR;
    }
  }
  return NANORQ_SYM_ADDED;
}

int nanorq_decoder_add_symbols(nanorq *rq, const uint32_t *tags, const void *data, size_t pitch, size_t n,
                               int *status, struct ioctx *io) {
  if (!rq || !tags || !data || pitch < rq->T) return -1;
  PF_T0;
  const uint8_t *rows = data;
  int added = 0, err = 0;
  size_t k = 0;
  while (k < n) {
    /* a run of symbols of one block: all of them are copied into consecutive input rows with
     * ONE copy from the caller's memory; the ones the block does not accept (duplicates,
     * symbols after completion) simply leave their row unused */
    if (tags[k] == NANORQ_TAG_NONE) { /* a hole at the start of a run */
      if (status) status[k] = NANORQ_SYM_IGN;
      k++;
      continue;
    }
    const uint8_t sbn = (tags[k] >> 24) & 0xff;
    size_t run = 1, last = 1; /* holes inside a run travel with it; holes at its end do not */
    while (k + run < n && (tags[k + run] == NANORQ_TAG_NONE || ((tags[k + run] >> 24) & 0xff) == sbn)) {
      run++;
      if (tags[k + run - 1] != NANORQ_TAG_NONE) last = run;
    }
    for (size_t q = last; q < run; q++)
      if (status) status[k + q] = NANORQ_SYM_IGN;
    const size_t skipped = run - last;
    run = last;
    struct block *b = get_block(rq, sbn);
    size_t take = run;
    if (b && b->mask && b->landed + take > b->in_cap) take = b->in_cap - b->landed; /* never more rows than the block has */
    bool copied = false;
    const uint32_t row0 = b ? b->landed : 0;
    for (size_t q = 0; q < run; q++) {
      if (tags[k + q] == NANORQ_TAG_NONE) {
        if (status) status[k + q] = NANORQ_SYM_IGN;
        continue;
      }
      const uint32_t esi = tags[k + q] & 0x00ffffff;
      int st = q < take ? classify(rq, b, esi) : NANORQ_SYM_ERR;
      if (st == NANORQ_SYM_ADDED) {
        if (!copied) { /* first accepted symbol of the run: queue the copy of the run's rows */
          if (!b->sv) { /* a small block without a context: its rows are collected by the CPU */
            for (size_t r = 0; r < take; r++)
              memcpy(b->lazy_stage + (size_t)(row0 + r) * b->pitch, rows + (k + r) * pitch, rq->T);
            copied = true;
            b->landed = row0 + (uint32_t)take;
          } else if (!flush_staged(b) || rqb_solver_upload_rows(b->sv, row0, (uint32_t)take, rows + k * pitch, pitch)) {
            st = NANORQ_SYM_ERR;
          } else {
            copied = true;
            b->landed = b->staged_lo = b->staged_hi = row0 + (uint32_t)take;
          }
        }
        if (st == NANORQ_SYM_ADDED) st = note_symbol(b, esi, row0 + (uint32_t)q);
        if (st == NANORQ_SYM_ADDED && esi < b->K && !deferred_output(rq, sbn, b, io))
          transfer_symbol(rq, sbn, esi, (uint8_t *)(uintptr_t)(rows + (k + q) * pitch), io, 1);
      }
      if (status) status[k + q] = st;
      added += st == NANORQ_SYM_ADDED;
      err |= st == NANORQ_SYM_ERR;
    }
    if (b && b->mask && b->gaps == 0 && copied) {
      if (b->out_mem) {
        if (!complete_block(rq, io, sbn, b)) err = 1;
      } else {
        rqb_copy_fence();
      }
    }
    k += run + skipped;
  }
  PF(RQB_PF_ADDS);
  return err ? -1 : added;
}

size_t nanorq_num_missing(nanorq *rq, uint8_t sbn) { /* :511-517 */
  struct block *b = get_block(rq, sbn);
  return b && b->mask ? b->gaps : 0;
}

size_t nanorq_num_repair(nanorq *rq, uint8_t sbn) { /* :519-525 */
  struct block *b = get_block(rq, sbn);
  return b ? b->nrep : 0;
}

/* deferred output: received source symbols are placed next to the recovered ones in the block
 * image on the device (emitted-symbol row = ESI) and the image goes to the caller's page-locked
 * memory with one copy */
static bool write_block_image(nanorq *rq, struct block *b, const uint32_t *have_esi, const uint32_t *have_row,
                              uint32_t n_have) {
  if (rqb_solver_copy_in_to_sym(b->sv, have_esi, have_row, n_have)) return false;
  const uint32_t full = (uint32_t)(b->out_bytes / rq->T);
  if (full && rqb_solver_fetch_rows(b->sv, 1, 0, full, b->out_mem, rq->T, 0)) return false;
  if (full < b->K && b->out_bytes > (size_t)full * rq->T) { /* the object's short last symbol */
    if (rqb_solver_fetch_syms(b->sv, full, 1, NULL, 0)) return false;
    memcpy(b->out_mem + (size_t)full * rq->T, rqb_solver_sym_mirror(b->sv) + (size_t)full * b->pitch,
           b->out_bytes - (size_t)full * rq->T);
  }
  return rqb_solver_sync(b->sv) == 0;
}

/* every source symbol has arrived and the output is deferred: hand the block back */
static bool complete_block(nanorq *rq, struct ioctx *io, uint8_t sbn, struct block *b) {
  (void)io;
  (void)sbn;
  if (b->written) return true;
  if (!flush_staged(b)) return false;
  uint32_t *esi = malloc(sizeof(uint32_t) * b->K);
  if (!esi) return false;
  for (uint32_t e = 0; e < b->K; e++) esi[e] = e;
  bool ok = write_block_image(rq, b, esi, b->src_row, b->K);
  free(esi);
  b->written = ok;
  return ok;
}

/* One block's repair in three steps, so that several blocks can share a launch:
 *   repair_prepare  fill_symbol_matrix_gaps + patch_precode_matrix (:527-565): request, host analysis,
 *                   program upload;  1 = ready to run, 0 = nothing to run (done or not decodable yet:
 *                   *result says which), -1 = the constraint matrix is singular / an error
 *   (launch)        rqb_solver_run for one block, rqb_solver_run_batch for several
 *   repair_finish   decode_repair_rows + write_repair_rows (:567-589) */
struct repair_job {
  struct block *b;
  uint8_t sbn;
  bool deferred;
  uint32_t *missing, *have; /* missing ESIs; deferred output: [K] ESIs received, then [K] their input rows */
  size_t nm, nh;
  rqb_solve_request req; /* valid between repair_request and repair_planned */
  uint32_t *isi, *in_row;
};

static void repair_job_free(struct repair_job *j) {
  free(j->missing);
  free(j->have);
  free(j->isi);
  free(j->in_row);
  j->missing = j->have = j->isi = j->in_row = NULL;
}

/* step 1a: the request (which symbol is where).  1 = built (j->req is valid until repair_job_free),
 * 0 = nothing to run (*result), -1 = error */
static int repair_request(nanorq *rq, struct ioctx *io, uint8_t sbn, struct repair_job *j, bool *result) {
  memset(j, 0, sizeof(*j));
  struct block *b = get_block(rq, sbn);
  *result = false;
  if (!b || !b->mask) return 0;
  if (b->gaps == 0) {
    *result = !b->out_mem || complete_block(rq, io, sbn, b);
    return 0;
  }
  if (b->nrep < b->gaps) return 0;
  const int Kp = rq->P.Kprime;
  const size_t gaps = b->gaps, overhead = b->nrep - gaps;
  const uint32_t pad = (uint32_t)Kp - b->K;
  const bool deferred = deferred_output(rq, sbn, b, io);
  /* the symbol bytes start moving to the GPU while the host analyses the matrix */
  PF_T0;
  if (!need_ctx(rq, b) || !flush_staged(b)) return -1;
  PF(RQB_PF_REP_UPLOAD);
  size_t nlt = (size_t)Kp + overhead;
  uint32_t *isi = malloc(sizeof(uint32_t) * nlt), *in_row = malloc(sizeof(uint32_t) * nlt);
  uint32_t *missing = malloc(sizeof(uint32_t) * gaps);
  uint32_t *have = deferred ? malloc(sizeof(uint32_t) * 2 * b->K) : NULL;
  if (!isi || !in_row || !missing || (deferred && !have)) {
    free(isi);
    free(in_row);
    free(missing);
    free(have);
    return -1;
  }
  size_t rep = 0, nm = 0, nh = 0;
  /* missing source rows take the repair symbols in arrival order, the rest become overhead rows */
  for (uint32_t e = 0; e < (uint32_t)Kp; e++) {
    if (e >= b->K) { /* padding symbol: known zero */
      isi[e] = e;
      in_row[e] = RQB_NO_ROW;
    } else if (mask_get(b, e)) {
      isi[e] = e;
      in_row[e] = b->src_row[e];
      if (have) {
        have[nh] = e;
        have[b->K + nh++] = b->src_row[e];
      }
    } else {
      isi[e] = b->rep_esi[rep] + pad;
      in_row[e] = b->rep_row[rep];
      rep++;
      missing[nm++] = e;
    }
  }
  for (size_t x = 0; x < overhead; x++, rep++) {
    isi[Kp + x] = b->rep_esi[rep] + pad;
    in_row[Kp + x] = b->rep_row[rep];
  }
  /* deferred output: the recovered symbol of ESI e is written to emitted-symbol row e */
  const rqb_solve_request req = {(int)overhead, isi, in_row, 0, (uint32_t)nm, missing, deferred ? missing : NULL};
  PF(RQB_PF_REP_REQUEST);
  j->req = req;
  j->isi = isi;
  j->in_row = in_row;
  j->b = b;
  j->sbn = sbn;
  j->deferred = deferred;
  j->missing = missing;
  j->have = have;
  j->nm = nm;
  j->nh = nh;
  return 1;
}

/* step 1b after rqb_solver_plan returned rc for j->req: 1 = ready to run, -1 = singular / error */
static int repair_planned(struct repair_job *j, int rc) {
  free(j->isi);
  free(j->in_row);
  j->isi = j->in_row = NULL;
  if (rc != 0) {
    rqb_solver_sync(j->b->sv);
    repair_job_free(j);
    return -1;
  }
  return 1;
}

static int repair_prepare(nanorq *rq, struct ioctx *io, uint8_t sbn, struct repair_job *j, bool *result) {
  int st = repair_request(rq, io, sbn, j, result);
  if (st <= 0) return st;
  return repair_planned(j, rqb_solver_plan(j->b->sv, &j->req)); /* charges repair.plan / .pages / .args itself */
}

static bool repair_finish(nanorq *rq, struct ioctx *io, struct repair_job *j) {
  struct block *b = j->b;
  bool ok = false;
  PF_T0;
  if (j->deferred) {
    ok = write_block_image(rq, b, j->have, j->have + b->K, (uint32_t)j->nh);
    PF(RQB_PF_REP_FETCH);
    if (ok) {
      for (size_t k = 0; k < j->nm; k++) mask_set(b, j->missing[k]);
      b->gaps = 0;
      b->written = true;
    }
  } else if (rqb_solver_fetch_syms(b->sv, 0, (uint32_t)j->nm, NULL, 0) == 0) {
    PF(RQB_PF_REP_FETCH);
    const uint8_t *sy = rqb_solver_sym_mirror(b->sv);
    for (size_t k = 0; k < j->nm; k++) {
      transfer_symbol(rq, j->sbn, j->missing[k], (uint8_t *)sy + k * b->pitch, io, 1);
      mask_set(b, j->missing[k]);
    }
    b->gaps = 0;
    ok = true;
    rqb_copy_fence();
    PF(RQB_PF_REP_WRITE);
  }
  if (!ok) rqb_solver_sync(b->sv);
  repair_job_free(j);
  return ok;
}

bool nanorq_repair_block(nanorq *rq, struct ioctx *io, uint8_t sbn) { /* :591-631 */
  struct repair_job j;
  bool result = false;
  int st = repair_prepare(rq, io, sbn, &j, &result);
  if (st <= 0) return st == 0 && result;
  PF_T0;
  if (rqb_solver_run(j.b->sv) != 0) {
    rqb_solver_sync(j.b->sv);
    repair_job_free(&j);
    return false;
  }
  PF(RQB_PF_REP_RUN);
  return repair_finish(rq, io, &j);
}

size_t nanorq_repair_blocks(nanorq *rq, struct ioctx *io, const uint8_t *sbns, size_t n, bool *ok) {
  /* every block is analysed first, then the solves of all blocks that live on one device run as ONE
   * kernel launch (gridDim.y = blocks), then the results are handed back block by block */
  if (!rq || !sbns) return 0;
  size_t good = 0;
  enum { MAXB = 255 };
  for (size_t at = 0; at < n;) {
    const size_t m = n - at < MAXB ? n - at : MAXB;
    struct repair_job *jobs = calloc(m, sizeof(*jobs));
    rqb_solver **sv = calloc(m, sizeof(*sv));
    int *state = calloc(m, sizeof(*state));
    bool *res = calloc(m, sizeof(*res)); /* outcome per entry (ok[] is optional) */
    if (!jobs || !sv || !state || !res) {
      free(jobs);
      free(sv);
      free(state);
      free(res);
      return good;
    }
#define SET_RESULT(k, v)              \
  do {                                \
    res[k] = (v);                     \
    if (ok) ok[at + (k)] = res[k];    \
  } while (0)
    /* requests first (cheap), then the analysis of all blocks side by side on the planning threads
     * (rqb_set_plan_threads; 1 = in this thread) */
    size_t np = 0;
    for (size_t k = 0; k < m; k++) {
      bool result = false;
      int first = -1; /* a block listed twice is repaired once; the later entries report the same outcome */
      for (size_t q = 0; q < k && first < 0; q++)
        if (sbns[at + q] == sbns[at + k]) first = (int)q;
      if (first >= 0) {
        state[k] = -2 - first;
        continue;
      }
      state[k] = repair_request(rq, io, sbns[at + k], &jobs[k], &result);
      if (state[k] == 0) SET_RESULT(k, result);
      if (state[k] < 0) SET_RESULT(k, false);
      if (state[k] == 0 && result) good++;
      if (state[k] == 1) np++;
    }
    if (np) {
      rqb_solve_request *reqs = malloc(sizeof(*reqs) * np);
      int *rcs = malloc(sizeof(int) * np);
      if (!reqs || !rcs) {
        free(reqs);
        free(rcs);
        for (size_t k = 0; k < m; k++)
          if (state[k] == 1) repair_job_free(&jobs[k]);
        free(jobs);
        free(sv);
        free(state);
        free(res);
        return good;
      }
      size_t q = 0;
      for (size_t k = 0; k < m; k++)
        if (state[k] == 1) {
          reqs[q] = jobs[k].req;
          sv[q++] = jobs[k].b->sv;
        }
      rqb_solver_plan_batch(sv, reqs, (int)np, rqb_get_plan_threads(), rcs);
      q = 0;
      for (size_t k = 0; k < m; k++)
        if (state[k] == 1) {
          state[k] = repair_planned(&jobs[k], rcs[q++]);
          if (state[k] < 0) SET_RESULT(k, false);
        }
      free(reqs);
      free(rcs);
    }
    /* launches: one per device among the prepared blocks */
    for (int dev = 0; dev < 64; dev++) {
      size_t cnt = 0;
      for (size_t k = 0; k < m; k++)
        if (state[k] == 1 && rqb_solver_device(jobs[k].b->sv) == dev) sv[cnt++] = jobs[k].b->sv;
      if (!cnt) continue;
      if (rqb_solver_run_batch(sv, (int)cnt) != 0)
        for (size_t k = 0; k < m; k++)
          if (state[k] == 1 && rqb_solver_device(jobs[k].b->sv) == dev) {
            rqb_solver_sync(jobs[k].b->sv);
            repair_job_free(&jobs[k]);
            state[k] = -1;
            SET_RESULT(k, false);
          }
    }
    for (size_t k = 0; k < m; k++) {
      if (state[k] <= -2) { /* listed before: entry -2 - state[k] has the outcome */
        SET_RESULT(k, res[(size_t)(-2 - state[k])]);
        good += res[k];
        continue;
      }
      if (state[k] != 1) continue;
      const bool r = repair_finish(rq, io, &jobs[k]);
      SET_RESULT(k, r);
      good += r;
    }
#undef SET_RESULT
    free(res);
    free(jobs);
    free(sv);
    free(state);
    at += m;
  }
  return good;
}
