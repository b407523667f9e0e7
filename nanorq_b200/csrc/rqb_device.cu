// rqb_device.cu -- sm_100a kernels for the nanorq hot path + the extern "C" shim.
//
// Kernels (reference function each one replaces):
//   rqb_solve_kernel   precode_matrix_apply_sched + precode_matrix_permute
//                      (lib/precode.c:3-32,379-389) and decode_row for the symbols
//                      requested with the solve (lib/nanorq.c:184-204).
//                      Column-sliced: one CTA owns a 128-byte column slice of every
//                      row of a source block and interprets the host-built program
//                      (rqb_program.h).  Rows stay in HBM/L2; a task is a gather of
//                      <= 8 row segments done by 8 lanes x 16 bytes (one 128-byte
//                      line per source, all loads in flight together).  Pages of
//                      the program are staged by TMA bulk copies (cp.async.bulk +
//                      mbarrier) into a shared-memory ring.
//   rqb_lt_kernel      decode_row / gen_tuple on demand (lib/nanorq.c:184-204,
//                      lib/tuple.c:21-43), tuples computed on the device.
//   rqb_rowops_kernel  oaxpy / oaddrow / oscal (deps/oblas/oblas_avx.c:43-114) as a
//                      batch of independent row ops streaming out of HBM with
//                      128-bit accesses.
//   rqb_gather_rows_kernel  precode_matrix_permute as an out-of-place gather.
//
// All arithmetic is GF(2)/GF(256) integer work (poly 0x11D); no tensor cores.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rfc6330_tables.h"
#include "rqb_device.h"
#include "rqb_program.h"

// ---------------------------------------------------------------- constants
__constant__ uint32_t c_rand_v[4][256];
__constant__ uint32_t c_degree_cdf[31];

#ifndef RQB_SOLVE_THREADS
#define RQB_SOLVE_THREADS 256
#endif
static constexpr int kSolveThreads = RQB_SOLVE_THREADS;
#ifndef RQB_SOLVE_MIN_CTAS
#define RQB_SOLVE_MIN_CTAS 4 /* 64 registers per thread: 4 CTAs = 1024 threads per SM (measured +18 % over 3) */
#endif
static constexpr int kSolveMinCtas = RQB_SOLVE_MIN_CTAS;       // CTAs per SM the register budget allows
// The column slice a CTA owns is a launch-time choice (template parameter kLanes = lanes of
// 16 bytes per task): 8 lanes = 128-byte slices is the default; 16 lanes = 256-byte slices for
// big batches (the thin levels of the triangular solve keep more of the CTA's lanes busy and
// every barrier covers twice the bytes); 4 lanes = 64-byte slices when only a few blocks are
// in flight (twice the CTAs for the fat levels).
static constexpr int kRingStages = 4; // 4 x 8 KiB pages in flight per CTA
static constexpr uint32_t kRingBytes = kRingStages * RQB_PAGE_BYTES;
static constexpr uint32_t kSolveSmem = kRingBytes + 128;       // ring + mbarriers

// ------------------------------------------------------------- GF(256) SWAR
// 4 packed field elements per 32-bit word.
__device__ __forceinline__ uint32_t xtime4(uint32_t x) {
  return ((x & 0x7f7f7f7fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x1du);
}
__device__ __forceinline__ uint32_t xtime1(uint32_t c) { // one element in the low byte
  return ((c << 1) ^ ((c & 0x80u) ? 0x11du : 0u)) & 0xffu;
}
// the eight products beta*2^k, k=0..7
struct BetaPlanes {
  uint32_t c[8];
};
__device__ __forceinline__ BetaPlanes beta_planes(uint32_t beta) {
  BetaPlanes p;
  p.c[0] = beta;
#pragma unroll
  for (int k = 1; k < 8; k++) p.c[k] = xtime1(p.c[k - 1]);
  return p;
}
// beta * x for 4 packed elements: sum over the bit planes of x
__device__ __forceinline__ uint32_t gfmul4(uint32_t x, const BetaPlanes &p) {
  uint32_t y = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) y ^= ((x >> k) & 0x01010101u) * p.c[k];
  return y;
}
__device__ __forceinline__ void xor4(uint4 &a, const uint4 &b) {
  a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w;
}

// --------------------------------------------------- mbarrier / TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy executed by the TMA unit, completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------- solve kernel
// grid = (ceil(width / 128), nblocks); block = kSolveThreads; dynamic smem = kSolveSmem.
// Thread t: lane q = t % 8 of task group t / 8; it owns bytes [col0 + 16q, +16) of
// every row its group touches.  Rows written in one level are read in later
// levels by other threads of the SAME CTA only (slices are disjoint), so the CTA
// barrier between levels is all the ordering the program needs.
// A row reference is the row number inside the block's arena (rqb_program.h):
// address = base + row * pitch, one IMAD.WIDE per source.
struct Rows {
  uint8_t *base; // arena, already offset to this thread's column
  uint32_t pitch;
  __device__ __forceinline__ uint8_t *at(uint32_t row) const {
    uint64_t p; // one IMAD.WIDE: base + row * pitch
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(row), "r"(pitch), "l"(base));
    return reinterpret_cast<uint8_t *>(p);
  }
  // 16 bytes of a row; rows are rewritten by other threads of the CTA between levels,
  // so the loads and stores keep their program order ("memory")
  __device__ __forceinline__ uint4 ld(uint32_t row) const {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(at(row))
                 : "memory");
    return v;
  }
  __device__ __forceinline__ void st(uint32_t row, const uint4 &v) const {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(at(row)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
};
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }

// The two rare task kinds live out of line so that the XOR gather (the hot path)
// gets the kernel's register budget to itself.
// SCAN: y = alpha*y ^ row[e_k]; row[dst+k] = y.  Loads run four entries ahead of the chain.
__device__ __noinline__ void task_scan(const Rows R, const uint32_t *sp, uint32_t nsrc, uint32_t dst) {
  uint4 y = make_uint4(0, 0, 0, 0);
  for (uint32_t k0 = 0; k0 < nsrc; k0 += 4) {
    const uint4 e = *reinterpret_cast<const uint4 *>(sp + k0);
    const uint32_t ev[4] = {e.x, e.y, e.z, e.w};
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      x[u] = make_uint4(0, 0, 0, 0);
      if (k0 + u < nsrc && ev[u] != RQB_REF_NONE) x[u] = R.ld(ev[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k0 + u < nsrc) {
        y.x = xtime4(y.x) ^ x[u].x; y.y = xtime4(y.y) ^ x[u].y;
        y.z = xtime4(y.z) ^ x[u].z; y.w = xtime4(y.w) ^ x[u].w;
        R.st(dst + k0 + u, y);
      }
    }
  }
}
// GF: row[dst] = XOR beta_k * row[src_k], nsrc <= 8; all loads first, then the multiplies.
__device__ __noinline__ void task_gf(const Rows R, const uint32_t *sp, uint32_t nsrc, uint32_t dst) {
  uint4 v[8];
#pragma unroll
  for (int u = 0; u < 8; u++) {
    v[u] = make_uint4(0, 0, 0, 0);
    if ((uint32_t)u < nsrc) v[u] = R.ld(sp[u] & RQB_REF_MASK);
  }
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < 8; u++) {
    if ((uint32_t)u < nsrc) {
      const BetaPlanes bp = beta_planes(sp[u] >> 24);
      acc.x ^= gfmul4(v[u].x, bp); acc.y ^= gfmul4(v[u].y, bp);
      acc.z ^= gfmul4(v[u].z, bp); acc.w ^= gfmul4(v[u].w, bp);
    }
  }
  R.st(dst, acc);
}

// TAB: row[dst] = row[src0] ^ XOR_j row[tab_base + 256*j + b_j] over the non-zero bytes b_j of
// one row of the bit matrix G (the back-substitution x = Y ^ G z through 8-bit XOR tables).
// Eight table rows are requested at a time; a zero byte selects the ZERO row, so the loads
// need no predicates.
__device__ __noinline__ void task_tab(const Rows R, const uint32_t *sp, uint32_t nbytes, uint32_t dst, uint32_t src0,
                                      uint32_t tab_base, uint32_t zero_row) {
  uint4 acc = R.ld(src0);
  for (uint32_t j0 = 0; j0 < nbytes; j0 += 8) {
    const uint2 w = *reinterpret_cast<const uint2 *>(sp + j0 / 4); // bytes j0 .. j0+7 (the list is zero-padded)
    uint32_t row[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t b = ((k < 4 ? w.x : w.y) >> (8 * (k & 3))) & 0xffu;
      row[k] = b ? tab_base + ((j0 + k) << 8) + b : zero_row;
    }
    uint4 v0 = R.ld(row[0]), v1 = R.ld(row[1]), v2 = R.ld(row[2]), v3 = R.ld(row[3]);
    uint4 v4 = R.ld(row[4]), v5 = R.ld(row[5]), v6 = R.ld(row[6]), v7 = R.ld(row[7]);
    acc.x ^= xor3(xor3(v0.x, v1.x, v2.x), xor3(v3.x, v4.x, v5.x), v6.x ^ v7.x);
    acc.y ^= xor3(xor3(v0.y, v1.y, v2.y), xor3(v3.y, v4.y, v5.y), v6.y ^ v7.y);
    acc.z ^= xor3(xor3(v0.z, v1.z, v2.z), xor3(v3.z, v4.z, v5.z), v6.z ^ v7.z);
    acc.w ^= xor3(xor3(v0.w, v1.w, v2.w), xor3(v3.w, v4.w, v5.w), v6.w ^ v7.w);
  }
  R.st(dst, acc);
}

template <int kLanes>
__global__ void __launch_bounds__(kSolveThreads, kSolveMinCtas)
rqb_solve_kernel(const rqb_solve_args *__restrict__ args_list) {
  constexpr int kLanesPerTask = kLanes;
  constexpr int kTaskGroups = kSolveThreads / kLanes;
  constexpr uint32_t kSliceBytes = 16u * kLanes;
  extern __shared__ __align__(128) uint8_t smem[];
  const rqb_solve_args &a = args_list[blockIdx.y];
  const uint32_t width = a.width;
  const uint32_t col0 = blockIdx.x * kSliceBytes;
  if (col0 >= width) return; // whole CTA leaves: no barrier is skipped by a subset
  uint8_t *ring = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kRingBytes);
  const int tid = threadIdx.x;
  const uint32_t n_pages = a.n_pages;
  const uint8_t *pages = a.pages;

  if (tid == 0) {
    for (int s = 0; s < kRingStages; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t first = n_pages < (uint32_t)kRingStages ? n_pages : (uint32_t)kRingStages;
    for (uint32_t s = 0; s < first; s++) {
      mbar_expect_tx(&bars[s], RQB_PAGE_BYTES);
      tma_bulk_g2s(ring + s * RQB_PAGE_BYTES, pages + (size_t)s * RQB_PAGE_BYTES, RQB_PAGE_BYTES, &bars[s]);
    }
  }
  const uint32_t col = col0 + (uint32_t)(tid % kLanesPerTask) * 16u;
  const bool active = col < width; // the last slice of a row may be narrower than 128 bytes
  const uint32_t grp = (uint32_t)tid / kLanesPerTask;
  Rows R;
  R.base = a.base + col;
  R.pitch = a.pitch;

  for (uint32_t pg = 0; pg < n_pages; pg++) {
    const uint32_t st = pg % kRingStages;
    mbar_wait(&bars[st], (pg / kRingStages) & 1u);
    const uint8_t *page = ring + st * RQB_PAGE_BYTES;
    const uint32_t n_levels = reinterpret_cast<const rqb_page_hdr *>(page)->n_levels;
    uint32_t off = sizeof(rqb_page_hdr);
    for (uint32_t lv = 0; lv < n_levels; lv++) {
      const uint4 lh = *reinterpret_cast<const uint4 *>(page + off); // n_tasks, next_off, tab_base, zero_row
      const uint4 *tasks = reinterpret_cast<const uint4 *>(page + off + sizeof(rqb_level_hdr));
      if (active) {
        for (uint32_t t = grp; t < lh.x; t += kTaskGroups) {
          const uint4 th = tasks[t]; // {src_off, dst, nsrc | kind<<16 | aux<<24, pad}
          const uint32_t nsrc = th.z & 0xffffu, kind = (th.z >> 16) & 0xffu;
          const uint32_t *sp = reinterpret_cast<const uint32_t *>(page + th.x);
          if (kind == RQB_T_XOR) {
            // the list holds exactly 4 or 8 rows (padded with the ZERO row): every load is
            // issued before the first value is consumed, no per-source predicates
            const uint4 i0 = *reinterpret_cast<const uint4 *>(sp);
            uint4 v0 = R.ld(i0.x), v1 = R.ld(i0.y), v2 = R.ld(i0.z), v3 = R.ld(i0.w);
            uint4 acc;
            if (nsrc > 4) {
              const uint4 i1 = *reinterpret_cast<const uint4 *>(sp + 4);
              uint4 v4 = R.ld(i1.x), v5 = R.ld(i1.y), v6 = R.ld(i1.z), v7 = R.ld(i1.w);
              acc.x = xor3(xor3(v0.x, v1.x, v2.x), xor3(v3.x, v4.x, v5.x), v6.x ^ v7.x);
              acc.y = xor3(xor3(v0.y, v1.y, v2.y), xor3(v3.y, v4.y, v5.y), v6.y ^ v7.y);
              acc.z = xor3(xor3(v0.z, v1.z, v2.z), xor3(v3.z, v4.z, v5.z), v6.z ^ v7.z);
              acc.w = xor3(xor3(v0.w, v1.w, v2.w), xor3(v3.w, v4.w, v5.w), v6.w ^ v7.w);
            } else {
              acc.x = xor3(v0.x, v1.x, v2.x) ^ v3.x;
              acc.y = xor3(v0.y, v1.y, v2.y) ^ v3.y;
              acc.z = xor3(v0.z, v1.z, v2.z) ^ v3.z;
              acc.w = xor3(v0.w, v1.w, v2.w) ^ v3.w;
            }
            R.st(th.y, acc);
          } else if (kind == RQB_T_TAB) {
            task_tab(R, sp, nsrc, th.y, th.w, lh.z, lh.w);
          } else if (kind == RQB_T_SCAN) {
            task_scan(R, sp, nsrc, th.y);
          } else {
            task_gf(R, sp, nsrc, th.y);
          }
        }
      }
      __syncthreads();
      off = lh.y;
    }
    if (n_levels == 0) __syncthreads();
    // every thread is past its last read of this stage: refill it
    if (tid == 0 && pg + kRingStages < n_pages) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[st], RQB_PAGE_BYTES);
      tma_bulk_g2s(ring + st * RQB_PAGE_BYTES, pages + (size_t)(pg + kRingStages) * RQB_PAGE_BYTES,
                   RQB_PAGE_BYTES, &bars[st]);
    }
  }
}

#ifdef RQB_TRACE
// debug builds only (-DRQB_TRACE): CTA (0,0) of the shared-memory kernel logs the clock after
// every page wait and every level barrier; rqb_dev_trace_fetch reads the log back
__device__ unsigned long long g_trace[16384];
__device__ unsigned g_trace_n;
// the event counter lives in a register of thread 0 (a counter in global memory would put an L2
// round trip into every event); stores are fire-and-forget
#define TRACE(tag, val)                                                                          \
  do {                                                                                           \
    if (trace_on && trace_n < 16384u)                                                            \
      g_trace[trace_n++] = ((unsigned long long)(tag) << 62) | ((unsigned long long)(val) << 40) | \
                           (clock64() & 0xFFFFFFFFFFull);                                         \
  } while (0)
#else
#define TRACE(tag, val) do { } while (0)
#endif

// ------------------------------------------------- shared-memory solve kernel
// The shared-memory flavour of the program (rqb_program.h): a CTA keeps its column slice
// (kLanes x 16 bytes) of every live row of the block in shared-memory SLOTS for the whole
// solve.  The received symbols are read from HBM once (LOAD tasks), every row operation of
// the elimination runs in place on the slots -- ~30-cycle shared-memory latency per
// dependency level instead of an L2/HBM round trip -- and only results go back to HBM, so
// DRAM traffic is the compulsory traffic.  One CTA per SM (the slots take most of the 227 KB),
// 512 threads = 512 / kLanes tasks at a time; program pages arrive through a 3-deep TMA ring.
// grid = (ceil(width / (16 kLanes)), nblocks); dynamic smem = ring + barriers + n_slots * 16 kLanes.
static constexpr int kSmemThreads = 512;
static constexpr int kSmemRingStages = RQB_SMEM_RING_STAGES; // the program stream is not the limit (a 6-deep ring of 4 KiB pages measured the same)
static constexpr uint32_t kSmemRingBytes = kSmemRingStages * RQB_PAGE_BYTES;
static constexpr uint32_t kSmemFixedBytes = kSmemRingBytes + 128; // ring + mbarriers, then the slots

template <int kLanes>
struct Slots {
  static constexpr uint32_t kSlice = 16u * kLanes;
  uint8_t *gbase;  // the block's HBM arena, already offset to this thread's column
  uint32_t pitch;
  uint32_t sbase;  // shared-space address of slot 0, already offset to this thread's lane
  __device__ __forceinline__ uint32_t saddr(uint32_t slot) const { return sbase + slot * kSlice; }
  __device__ __forceinline__ uint8_t *gaddr(uint32_t ref) const {
    uint64_t p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(ref & (RQB_REF_GLOBAL - 1u)), "r"(pitch), "l"(gbase));
    return reinterpret_cast<uint8_t *>(p);
  }
  __device__ __forceinline__ uint4 lds(uint32_t slot) const {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(saddr(slot))
                 : "memory");
    return v;
  }
  __device__ __forceinline__ void sts(uint32_t slot, const uint4 &v) const {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr(slot)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
  __device__ __forceinline__ uint4 ldg(uint32_t ref) const {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(gaddr(ref))
                 : "memory");
    return v;
  }
  __device__ __forceinline__ void stg(uint32_t ref, const uint4 &v) const {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr(ref)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
  // a mixed reference: a slot, or (bit 23) a row of the HBM arena
  __device__ __forceinline__ uint4 ld(uint32_t ref) const { return (ref & RQB_REF_GLOBAL) ? ldg(ref) : lds(ref); }
  __device__ __forceinline__ void st(uint32_t ref, const uint4 &v) const {
    if (ref & RQB_REF_GLOBAL) stg(ref, v); else sts(ref, v);
  }
};

template <int kLanes>
__device__ __noinline__ void smem_task_gf(const Slots<kLanes> R, const uint32_t *sp, uint32_t nsrc, uint32_t dst) {
  uint4 v[8];
#pragma unroll
  for (int u = 0; u < 8; u++) {
    v[u] = make_uint4(0, 0, 0, 0);
    if ((uint32_t)u < nsrc) v[u] = R.ld(sp[u] & RQB_REF_MASK);
  }
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < 8; u++) {
    if ((uint32_t)u < nsrc) {
      const BetaPlanes bp = beta_planes(sp[u] >> 24);
      acc.x ^= gfmul4(v[u].x, bp); acc.y ^= gfmul4(v[u].y, bp);
      acc.z ^= gfmul4(v[u].z, bp); acc.w ^= gfmul4(v[u].w, bp);
    }
  }
  R.st(dst, acc);
}

// SCAN2: y = alpha*y ^ slot[e_k]; the two HDPC accumulators of column k take y; y ends in slot `yend`
template <int kLanes>
__device__ __noinline__ void smem_task_scan2(const Slots<kLanes> R, const uint32_t *sp, uint32_t nsrc, uint32_t acc0,
                                             uint32_t yend, uint32_t aux) {
  const uint32_t H = aux & 31u;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (uint32_t h = 0; h < H; h++) R.sts(acc0 + h, zero);
  const uint32_t n_acc = (aux & 0x80u) ? nsrc - 1 : nsrc; // the last column of the scan carries no ones
  uint4 y = zero;
  for (uint32_t k0 = 0; k0 < nsrc; k0 += 4) {
    const uint4 e = *reinterpret_cast<const uint4 *>(sp + k0);
    const uint32_t ev[4] = {e.x, e.y, e.z, e.w};
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      x[u] = zero;
      if (k0 + u < nsrc && (ev[u] & RQB_REF_MASK) != RQB_REF_NONE) x[u] = R.lds(ev[u] & RQB_REF_MASK);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k0 + u < nsrc) {
        y.x = xtime4(y.x) ^ x[u].x; y.y = xtime4(y.y) ^ x[u].y;
        y.z = xtime4(y.z) ^ x[u].z; y.w = xtime4(y.w) ^ x[u].w;
        if (k0 + u < n_acc) {
          const uint32_t s1 = acc0 + ((ev[u] >> 24) & 15u), s2 = acc0 + (ev[u] >> 28);
          uint4 a1 = R.lds(s1), a2 = R.lds(s2);
          xor4(a1, y);
          xor4(a2, y);
          R.sts(s1, a1);
          R.sts(s2, a2);
        }
      }
    }
  }
  R.sts(yend, y);
}

// TAB, in place: slot[dst] ^= XOR_j slot[tab_base + (j << bits) + v_j] over the non-zero group
// values v_j of one row of G; the result also goes to the arena row `also` (an encoder's C row)
template <int kLanes>
__device__ __noinline__ void smem_task_tab(const Slots<kLanes> R, const uint32_t *sp, uint32_t ngroups, uint32_t dst,
                                           uint32_t also, uint32_t tab_base, uint32_t bits) {
  uint4 acc = R.lds(dst);
  for (uint32_t j0 = 0; j0 < ngroups; j0 += 8) {
    const uint2 w = *reinterpret_cast<const uint2 *>(sp + j0 / 4); // bytes j0 .. j0+7 (the list is zero-padded)
    uint32_t row[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t b = ((k < 4 ? w.x : w.y) >> (8 * (k & 3))) & 0xffu;
      row[k] = b ? tab_base + ((j0 + k) << bits) + b : 0u; // slot 0 is all zero
    }
    uint4 v0 = R.lds(row[0]), v1 = R.lds(row[1]), v2 = R.lds(row[2]), v3 = R.lds(row[3]);
    uint4 v4 = R.lds(row[4]), v5 = R.lds(row[5]), v6 = R.lds(row[6]), v7 = R.lds(row[7]);
    acc.x ^= xor3(xor3(v0.x, v1.x, v2.x), xor3(v3.x, v4.x, v5.x), v6.x ^ v7.x);
    acc.y ^= xor3(xor3(v0.y, v1.y, v2.y), xor3(v3.y, v4.y, v5.y), v6.y ^ v7.y);
    acc.z ^= xor3(xor3(v0.z, v1.z, v2.z), xor3(v3.z, v4.z, v5.z), v6.z ^ v7.z);
    acc.w ^= xor3(xor3(v0.w, v1.w, v2.w), xor3(v3.w, v4.w, v5.w), v6.w ^ v7.w);
  }
  R.sts(dst, acc);
  if (also != RQB_ROW_NONE) R.stg(also, acc);
}

template <int kLanes>
__global__ void __launch_bounds__(kSmemThreads, 1)
rqb_solve_smem_kernel(const rqb_solve_args *__restrict__ args_list) {
  constexpr int kTaskGroups = kSmemThreads / kLanes;
  constexpr uint32_t kSliceBytes = 16u * kLanes;
  extern __shared__ __align__(128) uint8_t smem[];
  const rqb_solve_args &a = args_list[blockIdx.y];
  const uint32_t width = a.width;
  const uint32_t col0 = blockIdx.x * kSliceBytes;
  if (col0 >= width) return; // whole CTA leaves: no barrier is skipped by a subset
  uint8_t *ring = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kSmemRingBytes);
  const int tid = threadIdx.x;
  const uint32_t n_pages = a.n_pages;
  const uint8_t *pages = a.pages;
  const uint32_t lane = (uint32_t)tid % kLanes, grp = (uint32_t)tid / kLanes;
  Slots<kLanes> R;
  R.gbase = a.base + col0 + lane * 16u;
  R.pitch = a.pitch;
  R.sbase = smem_u32(smem + kSmemFixedBytes) + lane * 16u;

  if (tid == 0) {
    for (int s = 0; s < kSmemRingStages; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (grp == 0) R.sts(0, make_uint4(0, 0, 0, 0)); // slot 0: the zero row
  __syncthreads();
  if (tid == 0) {
    const uint32_t first = n_pages < (uint32_t)kSmemRingStages ? n_pages : (uint32_t)kSmemRingStages;
    for (uint32_t s = 0; s < first; s++) {
      mbar_expect_tx(&bars[s], RQB_PAGE_BYTES);
      tma_bulk_g2s(ring + s * RQB_PAGE_BYTES, pages + (size_t)s * RQB_PAGE_BYTES, RQB_PAGE_BYTES, &bars[s]);
    }
  }
  const bool active = col0 + lane * 16u < width; // the last slice of a row may be narrower
#ifdef RQB_TRACE
  const bool trace_on = tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
  uint32_t trace_n = 0;
#endif
  TRACE(0, 0);

  for (uint32_t pg = 0; pg < n_pages; pg++) {
    const uint32_t st = pg % kSmemRingStages;
    mbar_wait(&bars[st], (pg / kSmemRingStages) & 1u);
    TRACE(1, pg);
    const uint8_t *page = ring + st * RQB_PAGE_BYTES;
    const uint32_t n_levels = reinterpret_cast<const rqb_page_hdr *>(page)->n_levels;
    uint32_t off = sizeof(rqb_page_hdr);
    for (uint32_t lv = 0; lv < n_levels; lv++) {
      const uint4 lh = *reinterpret_cast<const uint4 *>(page + off); // n_tasks, next_off, tab_base, zero_row
      const uint4 *tasks = reinterpret_cast<const uint4 *>(page + off + sizeof(rqb_level_hdr));
      if (active) {
        for (uint32_t t = grp; t < lh.x; t += kTaskGroups) {
          const uint4 th = tasks[t]; // {src_off, dst, nsrc | kind<<16 | aux<<24, pad}
          const uint32_t nsrc = th.z & 0xffffu, kind = (th.z >> 16) & 0xffu;
          const uint32_t *sp = reinterpret_cast<const uint32_t *>(page + th.x);
          if (kind == RQB_T_XOR) {
            // exactly 4 or 8 references (padded with slot 0): all loads issued before the first use.
            // aux bit 0: every reference is a slot -- the hot path of the thin levels: a level's latency
            // is the instruction count of one warp, so it carries no address arithmetic for HBM rows
            const uint4 i0 = *reinterpret_cast<const uint4 *>(sp);
            uint4 acc;
            if (th.z >> 24) {
              uint4 v0 = R.lds(i0.x), v1 = R.lds(i0.y), v2 = R.lds(i0.z), v3 = R.lds(i0.w);
              if (nsrc > 4) {
                const uint4 i1 = *reinterpret_cast<const uint4 *>(sp + 4);
                uint4 v4 = R.lds(i1.x), v5 = R.lds(i1.y), v6 = R.lds(i1.z), v7 = R.lds(i1.w);
                acc.x = xor3(xor3(v0.x, v1.x, v2.x), xor3(v3.x, v4.x, v5.x), v6.x ^ v7.x);
                acc.y = xor3(xor3(v0.y, v1.y, v2.y), xor3(v3.y, v4.y, v5.y), v6.y ^ v7.y);
                acc.z = xor3(xor3(v0.z, v1.z, v2.z), xor3(v3.z, v4.z, v5.z), v6.z ^ v7.z);
                acc.w = xor3(xor3(v0.w, v1.w, v2.w), xor3(v3.w, v4.w, v5.w), v6.w ^ v7.w);
              } else {
                acc.x = xor3(v0.x, v1.x, v2.x) ^ v3.x;
                acc.y = xor3(v0.y, v1.y, v2.y) ^ v3.y;
                acc.z = xor3(v0.z, v1.z, v2.z) ^ v3.z;
                acc.w = xor3(v0.w, v1.w, v2.w) ^ v3.w;
              }
              R.sts(th.y, acc);
            } else {
              uint4 v0 = R.ld(i0.x), v1 = R.ld(i0.y), v2 = R.ld(i0.z), v3 = R.ld(i0.w);
              if (nsrc > 4) {
                const uint4 i1 = *reinterpret_cast<const uint4 *>(sp + 4);
                uint4 v4 = R.ld(i1.x), v5 = R.ld(i1.y), v6 = R.ld(i1.z), v7 = R.ld(i1.w);
                acc.x = xor3(xor3(v0.x, v1.x, v2.x), xor3(v3.x, v4.x, v5.x), v6.x ^ v7.x);
                acc.y = xor3(xor3(v0.y, v1.y, v2.y), xor3(v3.y, v4.y, v5.y), v6.y ^ v7.y);
                acc.z = xor3(xor3(v0.z, v1.z, v2.z), xor3(v3.z, v4.z, v5.z), v6.z ^ v7.z);
                acc.w = xor3(xor3(v0.w, v1.w, v2.w), xor3(v3.w, v4.w, v5.w), v6.w ^ v7.w);
              } else {
                acc.x = xor3(v0.x, v1.x, v2.x) ^ v3.x;
                acc.y = xor3(v0.y, v1.y, v2.y) ^ v3.y;
                acc.z = xor3(v0.z, v1.z, v2.z) ^ v3.z;
                acc.w = xor3(v0.w, v1.w, v2.w) ^ v3.w;
              }
              R.st(th.y, acc);
            }
          } else if (kind == RQB_T_TAB) {
            smem_task_tab<kLanes>(R, sp, nsrc, th.y, th.w, lh.z, th.z >> 24);
          } else if (kind == RQB_T_LOAD) {
            // slots dst.. = arena rows pad..: four loads in flight
            for (uint32_t k = 0; k < nsrc; k += 4) {
              uint4 v[4];
#pragma unroll
              for (int u = 0; u < 4; u++)
                if (k + u < nsrc) v[u] = R.ldg(th.w + k + u);
#pragma unroll
              for (int u = 0; u < 4; u++)
                if (k + u < nsrc) R.sts(th.y + k + u, v[u]);
            }
          } else if (kind == RQB_T_SCAN2) {
            smem_task_scan2<kLanes>(R, sp, nsrc, th.y, th.w, th.z >> 24);
          } else {
            smem_task_gf<kLanes>(R, sp, nsrc, th.y);
          }
        }
      }
      TRACE(2, lh.x);
      __syncthreads();
      TRACE(3, lh.x);
      off = lh.y;
    }
    if (n_levels == 0) __syncthreads();
    // every thread is past its last read of this stage: refill it -- from the last warp, which has
    // tasks only in the fat levels (warp 0 is the one the thin levels wait for)
    if (tid == kSmemThreads - 1 && pg + kSmemRingStages < n_pages) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[st], RQB_PAGE_BYTES);
      tma_bulk_g2s(ring + st * RQB_PAGE_BYTES, pages + (size_t)(pg + kSmemRingStages) * RQB_PAGE_BYTES,
                   RQB_PAGE_BYTES, &bars[st]);
    }
  }
#ifdef RQB_TRACE
  if (trace_on) g_trace_n = trace_n;
#endif
}

// ------------------------------------------------------------- U-solve kernel
// The elimination of the u x u Schur system of one block (the reference's
// precode_matrix_solve_gf2 / precode_matrix_solve_gf256, lib/precode.c:264-315; steps 3d/3e of
// rqb_plan_build): Gauss-Jordan over GF(2) on the nb binary residual rows with the transformation
// tracked, then the H x nfree system of the HDPC rows over GF(256).  One CTA; the rows live in shared
// memory, every thread owns rows tid, tid + blockDim, ...  Pivot search = lowest row that is not a pivot
// yet and has the column's bit set: each thread offers its lowest such row, a warp-level min reduction
// (__reduce_min_sync) and a second one over the warps' results give the pivot.  The same rule as the
// host code, so both produce identical results.
struct rqb_usolve_hdr {
  int32_t nb, U, uw, nbw, H, sh_stride;
  // outputs
  int32_t status, nfree, rho, pad;
  int32_t qrow_of_f[16];
  uint8_t TQ[256];
};
// buffer layout after the header (all 16-byte aligned): Sb [nb*uw u64] | Tb [nb*nbw u64] | Sh [H*sh_stride] | pivrow [U i32]

__device__ __forceinline__ uint32_t gf_mul1(uint32_t a, uint32_t b) { // GF(256), poly 0x11D, shift-and-add
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    r ^= (b & 1u) ? a : 0u;
    a = xtime1(a);
    b >>= 1;
  }
  return r;
}
__device__ __forceinline__ uint32_t gf_inv1(uint32_t a) { // a^254
  uint32_t r = 1, p = a;
#pragma unroll
  for (int k = 1; k < 8; k++) { // 254 = 2 + 4 + ... + 128
    p = gf_mul1(p, p);
    r = gf_mul1(r, p);
  }
  return r;
}

__global__ void __launch_bounds__(1024, 1)
rqb_usolve_kernel(uint8_t *__restrict__ buf) {
  extern __shared__ __align__(16) uint8_t usm[];
  rqb_usolve_hdr *hdr = reinterpret_cast<rqb_usolve_hdr *>(buf);
  const int nb = hdr->nb, U = hdr->U, uw = hdr->uw, nbw = hdr->nbw, H = hdr->H, shs = hdr->sh_stride;
  const int rw = uw + nbw; // words per row in shared memory: [Schur bits | transformation bits]
  uint64_t *g_sb = reinterpret_cast<uint64_t *>(buf + sizeof(rqb_usolve_hdr));
  uint64_t *g_tb = g_sb + (size_t)nb * uw;
  const uint8_t *g_sh = reinterpret_cast<const uint8_t *>(g_tb + (size_t)nb * nbw);
  int32_t *g_piv = reinterpret_cast<int32_t *>(const_cast<uint8_t *>(g_sh) + (((size_t)H * shs + 15) & ~(size_t)15));
  uint64_t *rows = reinterpret_cast<uint64_t *>(usm);                 // [nb][rw]
  uint8_t *used = reinterpret_cast<uint8_t *>(rows + (size_t)nb * rw); // [nb]
  int32_t *piv = reinterpret_cast<int32_t *>(used + ((nb + 15) & ~15)); // [U]
  __shared__ int s_first[32];
  __shared__ int s_pr, s_nfree, s_rho;
  __shared__ uint8_t s_m[16][32]; // the HDPC system [Q | TQ], one column per lane
  __shared__ int s_free[16];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;

  for (int m = tid; m < nb; m += nthr) {
    for (int w = 0; w < uw; w++) rows[(size_t)m * rw + w] = g_sb[(size_t)m * uw + w];
    for (int w = 0; w < nbw; w++) rows[(size_t)m * rw + uw + w] = (w == (m >> 6)) ? (1ull << (m & 63)) : 0ull;
    used[m] = 0;
  }
  if (tid == 0) s_nfree = s_rho = 0;
  __syncthreads();

  for (int t = 0; t < U; t++) {
    int cand = 0x7fffffff;
    for (int m = tid; m < nb; m += nthr)
      if (!used[m] && ((rows[(size_t)m * rw + (t >> 6)] >> (t & 63)) & 1ull)) {
        cand = m;
        break;
      }
    const int wmin = __reduce_min_sync(0xffffffffu, cand); // warp-level reduction for the pivot search
    if (lane == 0) s_first[warp] = wmin;
    __syncthreads();
    if (warp == 0) {
      const int v = __reduce_min_sync(0xffffffffu, lane < nwarps ? s_first[lane] : 0x7fffffff);
      if (lane == 0) {
        s_pr = v;
        if (v == 0x7fffffff) {
          piv[t] = -1;
          if (s_nfree < 16) s_free[s_nfree] = t;
          s_nfree++;
        } else {
          piv[t] = v;
          used[v] = 1;
          s_rho++;
        }
      }
    }
    __syncthreads();
    const int pr = s_pr;
    if (pr == 0x7fffffff) {
      if (s_nfree > H) break; // more free columns than HDPC rows: singular (uniform: s_nfree is shared)
      continue;
    }
    const uint64_t *prow = rows + (size_t)pr * rw;
    for (int m = tid; m < nb; m += nthr)
      if (m != pr && ((rows[(size_t)m * rw + (t >> 6)] >> (t & 63)) & 1ull)) {
        uint64_t *r = rows + (size_t)m * rw;
        for (int w = 0; w < rw; w++) r[w] ^= prow[w];
      }
    __syncthreads();
  }
  const int nfree = s_nfree;
  if (nfree > H) {
    if (tid == 0) {
      hdr->status = 1;
      hdr->nfree = nfree;
      hdr->rho = s_rho;
    }
    return;
  }
  // HDPC rows: Q[h][f] = Sh[h][free f] ^ sum over pivot columns t of Sh[h][t] * (bit `free f` of t's pivot row)
  for (int e = tid; e < H * 32; e += nthr) {
    const int h = e >> 5, k = e & 31;
    uint32_t q = 0;
    if (k < nfree) {
      const int fc = s_free[k];
      const uint8_t *row = g_sh + (size_t)h * shs;
      q = row[fc];
      for (int t = 0; t < U; t++) {
        const uint32_t beta = row[t];
        if (beta && piv[t] >= 0 && ((rows[(size_t)piv[t] * rw + (fc >> 6)] >> (fc & 63)) & 1ull)) q ^= beta;
      }
    } else if (k < nfree + H) {
      q = (k - nfree == h) ? 1u : 0u; // the tracked transformation starts as the identity
    }
    s_m[h][k] = (uint8_t)q;
  }
  __syncthreads();
  // Gauss-Jordan over GF(256) on [Q | TQ]: warp 0, one column per lane
  if (warp == 0) {
    uint32_t usedmask = 0;
    int status = 0;
    for (int f = 0; f < nfree; f++) {
      int pr = -1;
      for (int h = 0; h < H; h++)
        if (!((usedmask >> h) & 1u) && s_m[h][f]) {
          pr = h;
          break;
        }
      if (pr < 0) {
        status = 1; // rank(A) < L
        break;
      }
      usedmask |= 1u << pr;
      if (lane == 0) hdr->qrow_of_f[f] = pr;
      const uint32_t inv = gf_inv1(s_m[pr][f]);
      uint32_t bcol[16];
      for (int h = 0; h < H; h++) bcol[h] = s_m[h][f]; // column f before it is updated
      __syncwarp();
      const uint32_t pk = gf_mul1(s_m[pr][lane], inv);
      s_m[pr][lane] = (uint8_t)pk;
      for (int h = 0; h < H; h++)
        if (h != pr && bcol[h]) s_m[h][lane] ^= (uint8_t)gf_mul1(bcol[h], pk);
      __syncwarp();
    }
    if (lane == 0) {
      hdr->status = status;
      hdr->nfree = nfree;
      hdr->rho = s_rho;
    }
    for (int h = 0; h < H; h++)
      if (lane < H) hdr->TQ[h * H + lane] = s_m[h][nfree + lane];
  }
  __syncthreads();
  for (int m = tid; m < nb; m += nthr) {
    for (int w = 0; w < uw; w++) g_sb[(size_t)m * uw + w] = rows[(size_t)m * rw + w];
    for (int w = 0; w < nbw; w++) g_tb[(size_t)m * nbw + w] = rows[(size_t)m * rw + uw + w];
  }
  for (int t = tid; t < U; t += nthr) g_piv[t] = piv[t];
}

// ---------------------------------------------------------------- LT kernel
// one CTA per output symbol; lane 0 expands Tuple[K', isi] into row indices.
__global__ void __launch_bounds__(128)
rqb_lt_kernel(rqb_params P, const uint8_t *__restrict__ c, uint32_t c_pitch,
              const uint32_t *__restrict__ isi, uint8_t *__restrict__ out, uint32_t out_pitch,
              uint32_t width) {
  __shared__ uint32_t idx[RQB_MAX_LT_DEGREE];
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = rqb_lt_indices(&P, c_rand_v, c_degree_cdf, isi[blockIdx.x], idx);
  __syncthreads();
  const int n = cnt;
  for (uint32_t v = threadIdx.x * 16; v < width; v += blockDim.x * 16) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int k = 0; k < n; k++) {
      const uint4 x = __ldg(reinterpret_cast<const uint4 *>(c + (size_t)idx[k] * c_pitch + v));
      acc.x ^= x.x; acc.y ^= x.y; acc.z ^= x.z; acc.w ^= x.w;
    }
    *reinterpret_cast<uint4 *>(out + (size_t)blockIdx.x * out_pitch + v) = acc;
  }
}

// ------------------------------------------------------------ row-op kernel
// flat index over (op, 16-byte chunk); reads src+dst, writes dst: 3*T bytes per axpy.
__global__ void __launch_bounds__(256)
rqb_rowops_kernel(uint8_t *__restrict__ D, size_t pitch, uint32_t vpr /* uint4 per row */,
                  const rqb_rowop *__restrict__ ops, uint32_t n) {
  const uint64_t total = (uint64_t)n * vpr;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t o = (uint32_t)(g / vpr), v = (uint32_t)(g - (uint64_t)o * vpr);
    const uint32_t beta = ops[o].beta, i = ops[o].i, j = ops[o].j;
    uint4 *dp = reinterpret_cast<uint4 *>(D + (size_t)i * pitch) + v;
    if (beta == 0) { // oscal: multiplier in j; u < 2 is a no-op (oblas_avx.c:94-95)
      const uint32_t u = j & 0xffu;
      if (u < 2) continue;
      const BetaPlanes bp = beta_planes(u);
      uint4 d = *dp;
      d.x = gfmul4(d.x, bp); d.y = gfmul4(d.y, bp); d.z = gfmul4(d.z, bp); d.w = gfmul4(d.w, bp);
      *dp = d;
    } else {
      const uint4 s = *(reinterpret_cast<const uint4 *>(D + (size_t)j * pitch) + v);
      uint4 d = *dp;
      if (beta == 1) {
        d.x ^= s.x; d.y ^= s.y; d.z ^= s.z; d.w ^= s.w;
      } else {
        const BetaPlanes bp = beta_planes(beta);
        d.x ^= gfmul4(s.x, bp); d.y ^= gfmul4(s.y, bp); d.z ^= gfmul4(s.z, bp); d.w ^= gfmul4(s.w, bp);
      }
      *dp = d;
    }
  }
}

__global__ void __launch_bounds__(256)
rqb_gather_rows_kernel(uint8_t *__restrict__ dst, size_t dpitch, const uint8_t *__restrict__ src,
                       size_t spitch, const uint32_t *__restrict__ map, uint32_t n, uint32_t vpr) {
  const uint64_t total = (uint64_t)n * vpr;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(g / vpr), v = (uint32_t)(g - (uint64_t)r * vpr);
    reinterpret_cast<uint4 *>(dst + (size_t)r * dpitch)[v] =
        __ldg(reinterpret_cast<const uint4 *>(src + (size_t)map[r] * spitch) + v);
  }
}

// row[pairs[2k+1]] = row[pairs[2k]] inside one arena
__global__ void __launch_bounds__(256)
rqb_copy_rows_kernel(uint8_t *__restrict__ base, size_t pitch, const uint32_t *__restrict__ pairs, uint32_t n,
                     uint32_t vpr) {
  const uint64_t total = (uint64_t)n * vpr;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(g / vpr), v = (uint32_t)(g - (uint64_t)r * vpr);
    reinterpret_cast<uint4 *>(base + (size_t)pairs[2 * r + 1] * pitch)[v] =
        __ldg(reinterpret_cast<const uint4 *>(base + (size_t)pairs[2 * r] * pitch) + v);
  }
}

// dst row k (dpitch apart) = src row k (spitch apart), `width` bytes each: re-pitches rows on the device
// so that host<->device copies are always linear (a 2-D DMA of ~1 KB rows reaches a fifth of the link)
template <typename V>
__global__ void __launch_bounds__(256)
rqb_repitch_kernel(uint8_t *__restrict__ dst, size_t dpitch, const uint8_t *__restrict__ src, size_t spitch,
                   uint32_t vpr /* V per row */, uint32_t n) {
  const uint64_t total = (uint64_t)n * vpr;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(g / vpr), v = (uint32_t)(g - (uint64_t)r * vpr);
    reinterpret_cast<V *>(dst + (size_t)r * dpitch)[v] = reinterpret_cast<const V *>(src + (size_t)r * spitch)[v];
  }
}

// ------------------------------------------------------------------- shim
static thread_local char g_err[256] = "";
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};
static std::atomic<int> g_consts_dev[64];
static std::atomic<int> g_default_dev{0};
static std::mutex g_cfg_mu;

static int fail(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
#define CK(x)                          \
  do {                                 \
    int _e = fail((x), #x);            \
    if (_e) return _e;                 \
  } while (0)

static int ensure_consts() {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 64 && g_consts_dev[dev].load(std::memory_order_acquire)) return 0;
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  CK(cudaMemcpyToSymbol(c_rand_v, rqb_rand_v, sizeof(rqb_rand_v)));
  CK(cudaMemcpyToSymbol(c_degree_cdf, rqb_degree_cdf, sizeof(rqb_degree_cdf)));
  if (dev < 64) g_consts_dev[dev].store(1, std::memory_order_release);
  return 0;
}

extern "C" {

const char *rqb_dev_last_error(void) { return g_err; }
unsigned long long rqb_dev_launch_count(void) { return g_launches.load(); }
void rqb_dev_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h) {
  if (h2d) *h2d = g_h2d_bytes.load();
  if (d2h) *d2h = g_d2h_bytes.load();
}
/* CUDA's current device is per thread (new threads start on device 0): the
 * library keeps a process-wide default that contexts are created on */
void rqb_dev_set_default(int dev) { g_default_dev.store(dev); }
int rqb_dev_default(void) { return g_default_dev.load(); }

int rqb_dev_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int rqb_dev_set(int dev) { CK(cudaSetDevice(dev)); return 0; }
int rqb_dev_get(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}
int rqb_dev_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
int rqb_dev_malloc(void **p, size_t bytes) { CK(cudaMalloc(p, bytes ? bytes : 16)); return 0; }
int rqb_dev_free(void *p) { CK(cudaFree(p)); return 0; }
int rqb_host_malloc(void **p, size_t bytes) { CK(cudaMallocHost(p, bytes ? bytes : 16)); return 0; }
int rqb_host_free(void *p) { CK(cudaFreeHost(p)); return 0; }
int rqb_stream_create(void **s) {
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = (void *)st;
  return 0;
}
int rqb_stream_destroy(void *s) { CK(cudaStreamDestroy((cudaStream_t)s)); return 0; }
// ---------------------------------------------------------------- waiting
// A host thread waits for its stream by watching a word in pinned host memory that a
// one-thread kernel at the end of the stream's queue sets (rqb_stream_wait_flag): no
// driver call while waiting, so a dozen waiting threads do not fight over the driver's
// lock with the threads that are launching work (polling cudaStreamQuery did: a 1.4 MB
// cudaMemcpyAsync took 3.5 ms at K=1024 with 20 threads).  The waiter spins briefly and
// then yields its core between looks, which is as prompt as spinning when every thread
// has a core and lets another worker run when there are more threads than cores.
// NANORQ_B200_WAIT = flag (default) | spin (cudaStreamSynchronize) | block (blocking-sync
// event: the thread sleeps; the wake-up costs ~0.5-1 ms on the B200 hosts, as long as a
// whole K=4096 solve).
__global__ void rqb_flag_kernel(volatile uint32_t *flag, uint32_t seq) {
  __threadfence_system();
  *flag = seq;
}

static int wait_mode() {
  static std::atomic<int> mode{-1};
  int m = mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char *e = getenv("NANORQ_B200_WAIT");
    m = 0;
    if (e && !strcmp(e, "spin")) m = 1;
    if (e && !strcmp(e, "block")) m = 2;
    mode.store(m, std::memory_order_relaxed);
  }
  return m;
}

int rqb_stream_sync(void *s) {
  static thread_local cudaEvent_t ev[64];
  int dev = 0;
  if (wait_mode() != 2 || cudaGetDevice(&dev) != cudaSuccess || dev >= 64) {
    CK(cudaStreamSynchronize((cudaStream_t)s));
    return 0;
  }
  if (!ev[dev]) CK(cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming));
  CK(cudaEventRecord(ev[dev], (cudaStream_t)s));
  CK(cudaEventSynchronize(ev[dev]));
  return 0;
}

// flag: a 32-bit word in pinned host memory owned by the caller; *seq is the last value
// handed out for it.  Returns when everything queued on the stream so far has run.
int rqb_stream_wait_flag(void *s, uint32_t *flag, uint32_t *seq) {
  if (wait_mode() != 0) return rqb_stream_sync(s);
  const uint32_t want = ++*seq;
  rqb_flag_kernel<<<1, 1, 0, (cudaStream_t)s>>>(flag, want); /* a signal, not counted in rqb_dev_launch_count */
  CK(cudaGetLastError());
  volatile uint32_t *f = flag;
  for (unsigned spins = 0; *f != want; spins++) {
    if (spins < 256) {
      for (int k = 0; k < 16; k++) __builtin_ia32_pause();
    } else {
      sched_yield();
      if ((spins & 0xfff) == 0) { /* a failed launch would never set the flag: look for errors now and then */
        cudaError_t q = cudaStreamQuery((cudaStream_t)s);
        if (q != cudaSuccess && q != cudaErrorNotReady) return fail(q, "stream failed while waiting");
      }
    }
  }
  return 0;
}
int rqb_dev_sync(void) { CK(cudaDeviceSynchronize()); return 0; }
int rqb_copy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  g_h2d_bytes += bytes;
  return 0;
}
int rqb_copy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  g_d2h_bytes += bytes;
  return 0;
}
int rqb_copy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int rqb_copy2d_h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                   void *stream) {
  CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  g_h2d_bytes += width * rows;
  return 0;
}
int rqb_copy2d_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                   void *stream) {
  CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  g_d2h_bytes += width * rows;
  return 0;
}
int rqb_dev_memset(void *p, int v, size_t bytes, void *stream) {
  CK(cudaMemsetAsync(p, v, bytes, (cudaStream_t)stream));
  return 0;
}
int rqb_event_create(void **e) {
  cudaEvent_t ev;
  CK(cudaEventCreate(&ev));
  *e = (void *)ev;
  return 0;
}
int rqb_event_create_sync(void **e) {
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  *e = (void *)ev;
  return 0;
}
int rqb_stream_wait_event(void *stream, void *event) {
  CK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return 0;
}
int rqb_dev_mem_info(size_t *free_bytes, size_t *total_bytes) {
  CK(cudaMemGetInfo(free_bytes, total_bytes));
  return 0;
}
int rqb_host_register(void *p, size_t bytes) {
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}
int rqb_host_unregister(void *p) {
  CK(cudaHostUnregister(p));
  return 0;
}
int rqb_event_destroy(void *e) { CK(cudaEventDestroy((cudaEvent_t)e)); return 0; }
int rqb_event_record(void *e, void *stream) { CK(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)stream)); return 0; }
int rqb_event_sync(void *e) { CK(cudaEventSynchronize((cudaEvent_t)e)); return 0; }
int rqb_event_elapsed_ms(void *start, void *stop, float *ms) {
  CK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

// slice width for a launch: NANORQ_B200_SLICE = 64 | 128 | 256 forces one (experiments); otherwise
// wide slices once 128-byte slices would fill the GPU about twice over, narrow ones when
// 128-byte slices would leave most SMs without a CTA
static int pick_slice_lanes(int nblocks, uint32_t max_width) {
  static std::atomic<int> forced{-1};
  int f = forced.load(std::memory_order_relaxed);
  if (f < 0) {
    const char *e = getenv("NANORQ_B200_SLICE");
    f = e ? atoi(e) / 16 : 0;
    if (f != 4 && f != 8 && f != 16) f = 0;
    forced.store(f, std::memory_order_relaxed);
  }
  if (f) return f;
  static std::atomic<int> sm_count{0}; /* one driver query per process (the GPUs of a box are alike) */
  int sms = sm_count.load(std::memory_order_relaxed);
  if (sms <= 0) {
    sms = rqb_dev_sm_count();
    if (sms <= 0) sms = 148;
    sm_count.store(sms, std::memory_order_relaxed);
  }
  /* measured on B200, K=4096, T=1280 (Gbit/s encode+decode at 30 / 59 / 118 blocks per launch):
   * 64-byte slices 787 / 1123 / 1153, 128-byte 966 / 1396 / 1452, 256-byte 723 / 1192 / 1636;
   * one block alone: 0.43 / 0.57 / 0.89 ms */
  const long ctas128 = (long)nblocks * ((max_width + 127) / 128);
  if (ctas128 >= 7L * sms) return 16;
  if (ctas128 * 2 <= sms) return 4;
  return 8;
}

int rqb_solve_slice_bytes(int nblocks, uint32_t max_width) { return 16 * pick_slice_lanes(nblocks, max_width); }

int rqb_launch_solve(const rqb_solve_args *args_dev, int nblocks, uint32_t max_width, void *stream) {
  if (nblocks <= 0 || max_width == 0) return 0;
  const int lanes = pick_slice_lanes(nblocks, max_width);
  const uint32_t slice = 16u * (uint32_t)lanes;
  dim3 grid((max_width + slice - 1) / slice, (unsigned)nblocks);
  if (lanes == 16)
    rqb_solve_kernel<16><<<grid, kSolveThreads, kSolveSmem, (cudaStream_t)stream>>>(args_dev);
  else if (lanes == 4)
    rqb_solve_kernel<4><<<grid, kSolveThreads, kSolveSmem, (cudaStream_t)stream>>>(args_dev);
  else
    rqb_solve_kernel<8><<<grid, kSolveThreads, kSolveSmem, (cudaStream_t)stream>>>(args_dev);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

/* shared-memory flavour: every block of the launch uses the same slice width (16, 32 or 64 bytes)
 * and at most max_slots slots */
int rqb_smem_slot_budget_bytes(void) { return (int)(227u * 1024u - kSmemFixedBytes); }

} // extern "C" (templates have C++ linkage)
template <int kLanes>
static int launch_smem(const rqb_solve_args *args_dev, dim3 grid, uint32_t smem_bytes, cudaStream_t stream) {
  static std::atomic<int> configured[64];
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    CK(cudaFuncSetAttribute(rqb_solve_smem_kernel<kLanes>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev < 64) configured[dev].store(1, std::memory_order_release);
  }
  rqb_solve_smem_kernel<kLanes><<<grid, kSmemThreads, smem_bytes, stream>>>(args_dev);
  return 0;
}
extern "C" {

int rqb_launch_solve_smem(const rqb_solve_args *args_dev, int nblocks, uint32_t max_width, uint32_t slice_bytes,
                          uint32_t max_slots, void *stream) {
  if (nblocks <= 0 || max_width == 0) return 0;
  const uint32_t smem_bytes = kSmemFixedBytes + max_slots * slice_bytes;
  if (smem_bytes > 227u * 1024u || (slice_bytes != 16 && slice_bytes != 32 && slice_bytes != 64)) {
    snprintf(g_err, sizeof(g_err), "rqb_launch_solve_smem: %u slots of %u bytes do not fit", max_slots, slice_bytes);
    return (int)cudaErrorInvalidValue;
  }
  dim3 grid((max_width + slice_bytes - 1) / slice_bytes, (unsigned)nblocks);
  int e;
  if (slice_bytes == 64)
    e = launch_smem<4>(args_dev, grid, smem_bytes, (cudaStream_t)stream);
  else if (slice_bytes == 32)
    e = launch_smem<2>(args_dev, grid, smem_bytes, (cudaStream_t)stream);
  else
    e = launch_smem<1>(args_dev, grid, smem_bytes, (cudaStream_t)stream);
  if (e) return e;
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

#ifdef RQB_TRACE
int rqb_dev_trace_fetch(unsigned long long *out, unsigned cap) {
  unsigned n = 0;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)));
  if (n > cap) n = cap;
  CK(cudaMemcpyFromSymbol(out, g_trace, (size_t)n * sizeof(unsigned long long)));
  return (int)n;
}
#endif

size_t rqb_usolve_buffer_bytes(int nb, int U, int uw, int nbw, int H, int sh_stride) {
  (void)U;
  return sizeof(rqb_usolve_hdr) + (size_t)nb * (size_t)(uw + nbw) * 8 + (((size_t)H * (size_t)sh_stride + 15) & ~(size_t)15) +
         (((size_t)U * 4 + 15) & ~(size_t)15);
}
size_t rqb_usolve_header_bytes(void) { return sizeof(rqb_usolve_hdr); }

/* buf_dev holds the header and the matrices (see rqb_usolve_kernel); returns cudaErrorInvalidValue when the
 * rows do not fit a CTA's shared memory (the host code runs then) */
int rqb_launch_usolve(uint8_t *buf_dev, int nb, int U, int uw, int nbw, void *stream) {
  const size_t smem = (size_t)nb * (size_t)(uw + nbw) * 8 + (size_t)((nb + 15) & ~15) + (size_t)U * 4 + 64;
  if (smem > 200u * 1024u) return (int)cudaErrorInvalidValue;
  static std::atomic<int> configured[64];
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    CK(cudaFuncSetAttribute(rqb_usolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) configured[dev].store(1, std::memory_order_release);
  }
  int threads = ((nb + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > 1024) threads = 1024;
  rqb_usolve_kernel<<<1, threads, smem, (cudaStream_t)stream>>>(buf_dev);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int rqb_launch_lt(const rqb_params *P, const uint8_t *c, uint32_t c_pitch, const uint32_t *isi_dev, uint32_t n,
                  uint8_t *out, uint32_t out_pitch, uint32_t width, void *stream) {
  if (n == 0) return 0;
  int e = ensure_consts();
  if (e) return e;
  rqb_lt_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(*P, c, c_pitch, isi_dev, out, out_pitch, width);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

static unsigned stream_grid(uint64_t total, int threads) {
  int sms = rqb_dev_sm_count();
  if (sms <= 0) sms = 148;
  uint64_t want = (total + threads - 1) / threads, cap = (uint64_t)sms * 16;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

int rqb_launch_rowops(uint8_t *D, size_t pitch, uint32_t width, const rqb_rowop *ops_dev, uint32_t n,
                      void *stream) {
  if (n == 0) return 0;
  const uint32_t vpr = width / 16;
  rqb_rowops_kernel<<<stream_grid((uint64_t)n * vpr, 256), 256, 0, (cudaStream_t)stream>>>(D, pitch, vpr, ops_dev, n);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int rqb_launch_copy_rows(uint8_t *base, size_t pitch, const uint32_t *pairs_dev, uint32_t n, uint32_t width,
                         void *stream) {
  if (n == 0) return 0;
  const uint32_t vpr = width / 16;
  rqb_copy_rows_kernel<<<stream_grid((uint64_t)n * vpr, 256), 256, 0, (cudaStream_t)stream>>>(base, pitch, pairs_dev,
                                                                                               n, vpr);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int rqb_launch_repitch(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, uint32_t width, uint32_t n,
                       void *stream) {
  if (n == 0 || width == 0) return 0;
  const uintptr_t al = (uintptr_t)dst | (uintptr_t)src | dpitch | spitch | width;
  cudaStream_t st = (cudaStream_t)stream;
  if (al % 16 == 0)
    rqb_repitch_kernel<uint4><<<stream_grid((uint64_t)n * (width / 16), 256), 256, 0, st>>>(dst, dpitch, src, spitch, width / 16, n);
  else if (al % 8 == 0)
    rqb_repitch_kernel<uint2><<<stream_grid((uint64_t)n * (width / 8), 256), 256, 0, st>>>(dst, dpitch, src, spitch, width / 8, n);
  else if (al % 4 == 0)
    rqb_repitch_kernel<uint32_t><<<stream_grid((uint64_t)n * (width / 4), 256), 256, 0, st>>>(dst, dpitch, src, spitch, width / 4, n);
  else
    rqb_repitch_kernel<uint8_t><<<stream_grid((uint64_t)n * width, 256), 256, 0, st>>>(dst, dpitch, src, spitch, width, n);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int rqb_launch_gather_rows(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch,
                           const uint32_t *map_dev, uint32_t n, uint32_t width, void *stream) {
  if (n == 0) return 0;
  const uint32_t vpr = width / 16;
  rqb_gather_rows_kernel<<<stream_grid((uint64_t)n * vpr, 256), 256, 0, (cudaStream_t)stream>>>(
      dst, dpitch, src, spitch, map_dev, n, vpr);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

} // extern "C"
