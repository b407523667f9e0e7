"""ctypes mirror of include/nanorq.h, include/io.h and include/rqb200.h.

Names and argument meaning follow the C headers (which follow the reference's
include/nanorq.h:16-83).  Nothing here computes on symbol bytes: every call goes
through libnanorq_b200.so, and the library fails loudly when no CUDA device is
visible (there is no CPU fallback)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NANORQ_B200_LIBDIR: load another build of the library (A/B measurements of two builds in one GPU session)
LIB_DIR = os.environ.get("NANORQ_B200_LIBDIR", _HERE)
LIB_PATH = os.path.join(LIB_DIR, "libnanorq_b200.so")

SYM_DUP, SYM_IGN, SYM_ADDED, SYM_ERR = 2, 1, 0, -1
NO_ROW = 0xFFFFFFFF
RQB_NEED_MORE = 1
RQB_E_NODEVICE = -100
RQB_E_TOOBIG = -101

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
vp = C.c_void_p


class BlockParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Kprime", "S", "H", "W", "L", "P", "P1", "U", "B", "J")]


class _SolveRequest(C.Structure):
    _fields_ = [("overhead", C.c_int), ("isi", u32p), ("in_row", u32p), ("want_c", C.c_int),
                ("n_out", C.c_uint32), ("out_isi", u32p), ("out_row", u32p)]


class SolverStats(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("i", "u", "nb", "rho", "nfree", "levels_fwd", "n_levels", "n_tasks", "n_pages")] + \
               [(n, C.c_size_t) for n in ("n_srcs", "n_gf_srcs", "n_horner", "nnz")] + \
               [(n, C.c_double) for n in ("t_matrix", "t_peel", "t_dense", "t_emit")] + \
               [("n_ws_rows", C.c_uint32), ("n_parts", C.c_int), ("slice_bytes", C.c_int), ("smem", C.c_int),
                ("n_slots", C.c_uint32), ("tab_bits", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PlanBlob(C.Structure):
    _fields_ = [("n_ws_rows", C.c_uint32), ("n_pages", C.c_uint32), ("page_bytes", C.c_uint32),
                ("row0", C.c_uint32 * 4), ("zero_row", C.c_uint32), ("n_rows", C.c_uint32),
                ("pages", u8p), ("stats", SolverStats), ("opaque", vp),
                ("smem", C.c_int), ("slice_bytes", C.c_uint32), ("n_slots", C.c_uint32), ("tab_bits", C.c_uint32)]


class Op(C.Structure):
    """binary-compatible with the reference's sched_op (include/sched.h:6-10)"""
    _fields_ = [("beta", C.c_uint8), ("i", C.c_uint32), ("j", C.c_uint32)]


OP_DTYPE = np.dtype({"names": ["beta", "i", "j"], "formats": [np.uint8, np.uint32, np.uint32],
                     "offsets": [0, 4, 8], "itemsize": 12})

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libnanorq_b200.so is not built: run `python -m nanorq_b200.build` "
            "(or __graft_entry__.build()); there is no pure-Python/CPU fallback")
    L = C.CDLL(LIB_PATH)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sz = C.c_size_t
    # nanorq.h
    sig("nanorq_encoder_new", vp, sz, C.c_uint16, C.c_uint8)
    sig("nanorq_encoder_new_ex", vp, sz, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint8)
    sig("nanorq_generate_symbols", C.c_bool, vp, C.c_uint8, vp)
    sig("nanorq_free", None, vp)
    sig("nanorq_oti_common", C.c_uint64, vp)
    sig("nanorq_oti_scheme_specific", C.c_uint32, vp)
    sig("nanorq_transfer_length", sz, vp)
    sig("nanorq_symbol_size", sz, vp)
    sig("nanorq_blocks", sz, vp)
    sig("nanorq_block_symbols", sz, vp, C.c_uint8)
    sig("nanorq_tag", C.c_uint32, C.c_uint8, C.c_uint32)
    sig("nanorq_max_blocks", sz, vp)
    sig("nanorq_precalculate", C.c_bool, vp)
    sig("nanorq_encode", sz, vp, vp, C.c_uint32, C.c_uint8, vp)
    sig("nanorq_encoder_cleanup", None, vp, C.c_uint8)
    sig("nanorq_encoder_reset", None, vp, C.c_uint8)
    sig("nanorq_decoder_new", vp, C.c_uint64, C.c_uint32)
    sig("nanorq_set_max_esi", C.c_bool, vp, C.c_uint32)
    sig("nanorq_decoder_add_symbol", C.c_int, vp, vp, C.c_uint32, vp)
    sig("nanorq_num_missing", sz, vp, C.c_uint8)
    sig("nanorq_num_repair", sz, vp, C.c_uint8)
    sig("nanorq_repair_block", C.c_bool, vp, vp, C.c_uint8)
    # nanorq_batch.h
    sig("ioctx_from_pinned_mem", vp, vp, sz, C.c_int)
    sig("nanorq_encode_range", sz, vp, C.c_uint8, C.c_uint32, C.c_uint32, vp, sz, vp)
    sig("nanorq_decoder_add_symbols", C.c_int, vp, u32p, vp, sz, sz, C.POINTER(C.c_int), vp)
    sig("nanorq_set_devices", C.c_int, vp, C.c_int)
    sig("nanorq_repair_blocks", sz, vp, vp, u8p, sz, C.POINTER(C.c_bool))
    # io.h
    sig("ioctx_from_file", vp, C.c_char_p, C.c_int)
    sig("ioctx_mmap_file", vp, C.c_char_p, C.c_int)
    sig("ioctx_from_mem", vp, vp, sz)
    # rqb200.h
    sig("rqb_last_error", C.c_char_p)
    sig("rqb_device_count", C.c_int)
    sig("rqb_set_device", C.c_int, C.c_int)
    sig("rqb_kernel_launches", C.c_ulonglong)
    sig("rqb_transfer_bytes", None, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong))
    sig("rqb_slow_path_counters", None, C.POINTER(C.c_ulonglong * 4))
    sig("rqb_host_profile", None, C.POINTER(C.c_double), C.c_int)
    sig("rqb_host_profile_name", C.c_char_p, C.c_int)
    sig("rqb_host_profile_reset", None)
    sig("rqb_solver_mark", C.c_int, vp, C.c_int)
    sig("rqb_solver_marked_ms", C.c_int, vp, C.POINTER(C.c_float))
    sig("rqb_block_params_init", C.c_int, C.c_int, C.POINTER(BlockParams))
    sig("rqb_lt_row_indices", C.c_int, C.c_int, C.c_uint32, u32p)
    sig("rqb_solver_create", C.c_int, C.POINTER(vp), C.c_int, sz, C.c_uint32, C.c_uint32)
    sig("rqb_solver_create_ex", C.c_int, C.POINTER(vp), C.c_int, C.c_int, sz, C.c_uint32, C.c_uint32)
    sig("rqb_solver_destroy", None, vp)
    sig("rqb_release_cached", None)
    sig("rqb_set_cache_limit", None, sz)
    sig("rqb_cache_stats", None, C.POINTER(sz), C.POINTER(sz))
    sig("rqb_device_mem_info", C.c_int, C.POINTER(sz), C.POINTER(sz))
    sig("rqb_solver_staging", vp, vp)
    sig("rqb_solver_pitch", sz, vp)
    sig("rqb_solver_upload", C.c_int, vp, C.c_uint32, C.c_uint32)
    sig("rqb_solver_plan", C.c_int, vp, C.POINTER(_SolveRequest))
    sig("rqb_solver_plan_encode", C.c_int, vp, C.c_int, C.c_uint32)
    sig("rqb_solver_plan_batch", C.c_int, C.POINTER(vp), C.POINTER(_SolveRequest), C.c_int, C.c_int, C.POINTER(C.c_int))
    sig("rqb_set_plan_threads", None, C.c_int)
    sig("rqb_get_plan_threads", C.c_int)
    sig("rqb_solver_run", C.c_int, vp)
    sig("rqb_solver_emit", C.c_int, vp, u32p, C.c_uint32)
    sig("rqb_solver_sync", C.c_int, vp)
    sig("rqb_solver_fetch_syms", C.c_int, vp, C.c_uint32, C.c_uint32, vp, sz)
    sig("rqb_solver_fetch_syms_async", C.c_int, vp, C.c_uint32, C.c_uint32)
    sig("rqb_solver_fetch_c", C.c_int, vp, C.c_uint32, C.c_uint32, vp, sz)
    sig("rqb_solver_sym_mirror", vp, vp)
    sig("rqb_solver_last_kernel_ms", C.c_int, vp, C.POINTER(C.c_float))
    sig("rqb_solver_set_timing", None, vp, C.c_int)
    sig("rqb_solver_get_stats", C.c_int, vp, C.POINTER(SolverStats))
    sig("rqb_batch_slice_bytes", C.c_int, C.c_int, sz)
    sig("rqb_solver_run_batch", C.c_int, C.POINTER(vp), C.c_int)
    sig("rqb_solver_run_batch_on", C.c_int, C.POINTER(vp), C.c_int, vp)
    sig("rqb_solver_create_on", C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, sz, C.c_uint32, C.c_uint32)
    sig("rqb_solver_device", C.c_int, vp)
    sig("rqb_solver_set_flavour", None, vp, C.c_int)
    sig("rqb_solver_upload_rows", C.c_int, vp, C.c_uint32, C.c_uint32, vp, sz)
    sig("rqb_solver_fetch_rows", C.c_int, vp, C.c_int, C.c_uint32, C.c_uint32, vp, sz, C.c_int)
    sig("rqb_solver_copy_in_to_sym", C.c_int, vp, u32p, u32p, C.c_uint32)
    sig("rqb_set_usolve_mode", None, C.c_int)
    sig("rqb_host_alloc", vp, sz)
    sig("rqb_host_release", None, vp)
    sig("rqb_host_pin", C.c_int, vp, sz)
    sig("rqb_host_unpin", C.c_int, vp)
    sig("rqb_plan_blob_build", C.c_int, C.c_int, C.POINTER(_SolveRequest), C.POINTER(PlanBlob))
    sig("rqb_plan_blob_build_ex", C.c_int, C.c_int, C.POINTER(_SolveRequest), C.c_uint32, C.POINTER(PlanBlob))
    sig("rqb_smem_budget", C.c_uint32)
    sig("rqb_plan_blob_free", None, C.POINTER(PlanBlob))
    sig("rqb_matrix_create", C.c_int, C.POINTER(vp), sz, sz)
    sig("rqb_matrix_destroy", None, vp)
    sig("rqb_matrix_pitch", sz, vp)
    sig("rqb_matrix_upload", C.c_int, vp, sz, sz, vp, sz)
    sig("rqb_matrix_download", C.c_int, vp, sz, sz, vp, sz)
    sig("rqb_matrix_fill_random", C.c_int, vp, C.c_uint64)
    sig("rqb_rowops_apply", C.c_int, vp, vp, sz)
    sig("rqb_ops_upload", C.c_int, C.POINTER(vp), vp, sz)
    sig("rqb_ops_free", None, vp)
    sig("rqb_rowops_apply_dev", C.c_int, vp, vp, C.c_int, C.POINTER(C.c_float))
    sig("rqb_schedule_plan_blob", C.c_int, sz, vp, sz, C.c_long, C.c_long, C.POINTER(C.c_int), sz,
        C.POINTER(C.c_int), sz, C.POINTER(PlanBlob))
    for name in ("rqb_schedule_replay", "rqb_schedule_replay_stepwise"):
        sig(name, C.c_int, vp, vp, sz, C.c_long, C.c_long, C.POINTER(C.c_int), sz,
            C.POINTER(C.c_int), sz, C.POINTER(C.c_float))
    _lib = L
    return L


# every symbol include/*.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = [
    "nanorq_encoder_new", "nanorq_encoder_new_ex", "nanorq_generate_symbols", "nanorq_free",
    "nanorq_oti_common", "nanorq_oti_scheme_specific", "nanorq_transfer_length", "nanorq_symbol_size",
    "nanorq_blocks", "nanorq_block_symbols", "nanorq_tag", "nanorq_max_blocks", "nanorq_precalculate",
    "nanorq_encode", "nanorq_encoder_cleanup", "nanorq_encoder_reset", "nanorq_decoder_new",
    "nanorq_set_max_esi", "nanorq_decoder_add_symbol", "nanorq_num_missing", "nanorq_num_repair",
    "nanorq_repair_block", "ioctx_from_file", "ioctx_mmap_file", "ioctx_from_mem",
    "rqb_last_error", "rqb_device_count", "rqb_set_device", "rqb_kernel_launches", "rqb_transfer_bytes",
    "rqb_solver_mark", "rqb_solver_marked_ms", "rqb_solver_run_batch_on",
    "rqb_host_profile", "rqb_host_profile_name", "rqb_host_profile_reset", "rqb_slow_path_counters",
    "rqb_block_params_init", "rqb_lt_row_indices", "rqb_solver_create", "rqb_solver_create_ex",
    "rqb_solver_destroy", "rqb_release_cached", "rqb_solver_staging", "rqb_solver_pitch", "rqb_solver_upload",
    "rqb_solver_plan", "rqb_solver_plan_encode", "rqb_solver_run", "rqb_solver_emit",
    "rqb_solver_sync", "rqb_solver_fetch_syms", "rqb_solver_fetch_c", "rqb_solver_fetch_syms_async", "rqb_solver_sym_mirror",
    "rqb_solver_last_kernel_ms", "rqb_solver_set_timing", "rqb_solver_get_stats", "rqb_solver_run_batch", "rqb_batch_slice_bytes", "rqb_plan_blob_build",
    "rqb_plan_blob_free", "rqb_matrix_create", "rqb_matrix_destroy", "rqb_matrix_pitch",
    "rqb_matrix_upload", "rqb_matrix_download", "rqb_matrix_fill_random", "rqb_rowops_apply",
    "rqb_ops_upload", "rqb_ops_free", "rqb_rowops_apply_dev", "rqb_schedule_replay",
    "rqb_schedule_replay_stepwise", "rqb_schedule_plan_blob",
    "rqb_set_cache_limit", "rqb_cache_stats", "rqb_device_mem_info", "rqb_plan_blob_build_ex", "rqb_smem_budget",
    "ioctx_from_pinned_mem", "nanorq_repair_blocks", "nanorq_encode_range", "nanorq_decoder_add_symbols", "nanorq_set_devices",
    "rqb_solver_create_on", "rqb_solver_device", "rqb_solver_set_flavour", "rqb_solver_upload_rows",
    "rqb_solver_fetch_rows", "rqb_solver_copy_in_to_sym", "rqb_host_alloc", "rqb_host_release", "rqb_host_pin", "rqb_set_usolve_mode",
    "rqb_solver_plan_batch", "rqb_set_plan_threads", "rqb_get_plan_threads",
    "rqb_host_unpin",
]


def last_error():
    return lib().rqb_last_error().decode()


def device_count():
    return lib().rqb_device_count()


def set_device(dev):
    _check(lib().rqb_set_device(dev), "rqb_set_device")


def kernel_launches():
    return int(lib().rqb_kernel_launches())


def transfer_bytes():
    h, d = C.c_ulonglong(), C.c_ulonglong()
    lib().rqb_transfer_bytes(C.byref(h), C.byref(d))
    return int(h.value), int(d.value)


def host_profile(reset=False):
    """seconds per slot of the nanorq.h layer, summed over threads (NANORQ_B200_PROFILE=1)"""
    out = (C.c_double * 32)()
    lib().rqb_host_profile(out, 32)
    names = []
    while lib().rqb_host_profile_name(len(names)):
        names.append(lib().rqb_host_profile_name(len(names)).decode())
    if reset:
        lib().rqb_host_profile_reset()
    return {n: out[k] for k, n in enumerate(names)}


def release_cached():
    lib().rqb_release_cached()


def set_cache_limit(nbytes):
    lib().rqb_set_cache_limit(int(nbytes))


def cache_stats():
    """-> (bytes parked in recycled contexts and pooled buffers, limit)"""
    a, b = C.c_size_t(), C.c_size_t()
    lib().rqb_cache_stats(C.byref(a), C.byref(b))
    return a.value, b.value


def device_mem_info():
    """-> (free, total) bytes of the library's default device"""
    a, b = C.c_size_t(), C.c_size_t()
    _check(lib().rqb_device_mem_info(C.byref(a), C.byref(b)), "rqb_device_mem_info")
    return a.value, b.value


def slow_path_counters():
    """{pinned allocations, device allocations, arena regrowths, contexts created} so far"""
    out = (C.c_ulonglong * 4)()
    lib().rqb_slow_path_counters(C.byref(out))
    return dict(zip(("alloc_pinned", "alloc_device", "arena_regrow", "contexts_new"), [int(x) for x in out]))


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, last_error()))


def block_params(K):
    p = BlockParams()
    if lib().rqb_block_params_init(K, C.byref(p)) != 0:
        raise ValueError("bad K %r" % (K,))
    return p


def lt_row_indices(K, isi):
    out = (C.c_uint32 * 40)()
    n = lib().rqb_lt_row_indices(K, isi, out)
    return list(out[:n])


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class SolveRequest:
    """rqb_solve_request: which LT rows exist, where their bytes are, what to emit."""

    def __init__(self, isi, in_row, overhead=0, want_c=True, out_isi=(), out_row=None):
        self.isi = _u32(isi)
        self.in_row = _u32(in_row)
        self.out_isi = _u32(out_isi)
        self.out_row = None if out_row is None else _u32(out_row)
        assert len(self.isi) == len(self.in_row)
        assert self.out_row is None or len(self.out_row) == len(self.out_isi)
        self.c = _SolveRequest(int(overhead), self.isi.ctypes.data_as(u32p), self.in_row.ctypes.data_as(u32p),
                               1 if want_c else 0, len(self.out_isi), self.out_isi.ctypes.data_as(u32p),
                               None if self.out_row is None else self.out_row.ctypes.data_as(u32p))

    @staticmethod
    def for_encoder(K, want_c=True, out_isi=()):
        p = block_params(K)
        k = np.arange(p.Kprime, dtype=np.uint32)
        return SolveRequest(k, np.where(k < K, k, NO_ROW), 0, want_c, out_isi)

    @staticmethod
    def for_decoder(K, esis, K_params=None, want_c=True):
        """Row placement of nanorq_decoder_add_symbol / nanorq_repair_block
        (reference lib/nanorq.c:478-509,527-565) for symbols that arrived in the
        order `esis` and were staged in that order (staging row = arrival index).
        Returns (request, missing_esis) or (None, missing) when too few symbols."""
        p = block_params(K_params or K)
        pad = p.Kprime - K
        have, reps, seen = {}, [], set()
        for k, e in enumerate(esis):
            e = int(e)
            if len(have) == K:
                break
            if e < K:
                have.setdefault(e, k)
            elif e not in seen:
                seen.add(e)
                reps.append((e, k))
        missing = [e for e in range(K) if e not in have]
        if len(reps) < len(missing):
            return None, missing
        oh = len(reps) - len(missing)
        isi = np.arange(p.Kprime + oh, dtype=np.uint32)
        in_row = np.full(p.Kprime + oh, NO_ROW, dtype=np.uint32)
        for e, k in have.items():
            in_row[e] = k
        for g, (e, k) in zip(missing, reps):
            isi[g] = e + pad
            in_row[g] = k
        for x, (e, k) in enumerate(reps[len(missing):]):
            isi[p.Kprime + x] = e + pad
            in_row[p.Kprime + x] = k
        # want_c=False is what nanorq_repair_block asks for: only the missing symbols come back
        return SolveRequest(isi, in_row, oh, want_c, missing), missing


def smem_budget():
    return lib().rqb_smem_budget()


def plan_blob(K_params, req, smem=0):
    """Host-only: build the device program and return (rc, dict) without a GPU.
    smem: bytes of shared memory for row slots (0 = HBM flavour, True = the kernel's budget)."""
    b = PlanBlob()
    if smem is True:
        smem = smem_budget()
    rc = lib().rqb_plan_blob_build_ex(K_params, C.byref(req.c), int(smem), C.byref(b))
    if rc != 0:
        return rc, None
    out = {
        "n_ws_rows": b.n_ws_rows, "n_pages": b.n_pages, "page_bytes": b.page_bytes,
        "row0": list(b.row0), "zero_row": b.zero_row, "n_rows": b.n_rows,
        "pages": np.ctypeslib.as_array(b.pages, (b.n_pages * b.page_bytes,)).copy(),
        "stats": b.stats.as_dict(),
        "smem": b.smem, "slice_bytes": b.slice_bytes, "n_slots": b.n_slots, "tab_bits": b.tab_bits,
    }
    lib().rqb_plan_blob_free(C.byref(b))
    return 0, out


def schedule_plan_blob(nrows, ops, mark0, mark1, di, c):
    """Host-only: the one-launch program of rqb_schedule_replay for a reference-format schedule."""
    assert ops.dtype == OP_DTYPE
    di = np.ascontiguousarray(di, dtype=np.int32)
    c = np.ascontiguousarray(c, dtype=np.int32)
    b = PlanBlob()
    rc = lib().rqb_schedule_plan_blob(nrows, ops.ctypes.data, len(ops), mark0, mark1,
                                      di.ctypes.data_as(C.POINTER(C.c_int)), len(di),
                                      c.ctypes.data_as(C.POINTER(C.c_int)), len(c), C.byref(b))
    if rc != 0:
        return rc, None
    out = {"n_ws_rows": b.n_ws_rows, "n_pages": b.n_pages, "page_bytes": b.page_bytes, "row0": list(b.row0),
           "zero_row": b.zero_row, "n_rows": b.n_rows,
           "pages": np.ctypeslib.as_array(b.pages, (b.n_pages * b.page_bytes,)).copy(), "stats": b.stats.as_dict()}
    lib().rqb_plan_blob_free(C.byref(b))
    return 0, out


class MemIO:
    """ioctx_from_mem over a numpy uint8 array (reference lib/io.c:139-157)."""

    ptr = None

    def __init__(self, arr):
        assert arr.dtype == np.uint8 and arr.flags.c_contiguous
        self.arr = arr
        self.ptr = lib().ioctx_from_mem(arr.ctypes.data, arr.size)

    def close(self):
        if self.ptr:
            destroy = C.cast(self.ptr, C.POINTER(_IoCtx)).contents.destroy
            destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


class PinnedBuffer:
    """Page-locked host memory from rqb_host_alloc as a numpy uint8 array (`.arr`): symbol buffers
    and payloads the GPU's copy engines reach without a staging copy."""

    ptr = None

    def __init__(self, nbytes):
        self.ptr = lib().rqb_host_alloc(nbytes)
        if not self.ptr:
            raise RuntimeError("rqb_host_alloc failed: " + last_error())
        self.arr = np.frombuffer((C.c_uint8 * nbytes).from_address(self.ptr), dtype=np.uint8)

    def close(self):
        if self.ptr:
            self.arr = None
            lib().rqb_host_release(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


class PinnedMemIO(MemIO):
    """ioctx_from_pinned_mem (nanorq_batch.h) over a numpy uint8 array; already_pinned=False page-locks
    it for the lifetime of the context."""

    def __init__(self, arr, already_pinned=True):
        assert arr.dtype == np.uint8 and arr.flags.c_contiguous
        self.arr = arr
        self.ptr = lib().ioctx_from_pinned_mem(arr.ctypes.data, arr.size, 1 if already_pinned else 0)
        if not self.ptr:
            raise RuntimeError("ioctx_from_pinned_mem failed: " + last_error())


class FileIO:
    """ioctx_from_file / ioctx_mmap_file (reference lib/io.c:54,338); mode 1 = read (encoder),
    0 = create (decoder)."""
    ptr = None

    def __init__(self, path, mode, mmap=False):
        fn = lib().ioctx_mmap_file if mmap else lib().ioctx_from_file
        self.ptr = fn(os.fsencode(path), mode)
        if not self.ptr:
            raise OSError("cannot open %r" % (path,))

    def close(self):
        if self.ptr:
            C.cast(self.ptr, C.POINTER(_IoCtx)).contents.destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


class _IoCtx(C.Structure):
    _fields_ = [("read", vp), ("write", vp), ("seek", vp), ("size", vp), ("tell", vp),
                ("destroy", C.CFUNCTYPE(None, vp)), ("seekable", C.c_bool), ("writable", C.c_bool)]


class _Codec:
    h = None

    def __init__(self, h):
        if not h:
            raise ValueError("nanorq constructor returned NULL")
        self.h = h

    def close(self):
        if self.h:
            lib().nanorq_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def oti_common(self):
        return lib().nanorq_oti_common(self.h)

    def oti_scheme_specific(self):
        return lib().nanorq_oti_scheme_specific(self.h)

    def transfer_length(self):
        return lib().nanorq_transfer_length(self.h)

    def symbol_size(self):
        return lib().nanorq_symbol_size(self.h)

    def blocks(self):
        return lib().nanorq_blocks(self.h)

    def block_symbols(self, sbn):
        return lib().nanorq_block_symbols(self.h, sbn)

    def encoder_cleanup(self, sbn):
        lib().nanorq_encoder_cleanup(self.h, sbn)

    def encoder_reset(self, sbn):
        lib().nanorq_encoder_reset(self.h, sbn)


class Encoder(_Codec):
    def __init__(self, length, T, K=0, Z=0, Al=8):
        super().__init__(lib().nanorq_encoder_new_ex(length, T, K, Z, Al))

    def precalculate(self):
        return lib().nanorq_precalculate(self.h)

    def generate_symbols(self, sbn, io):
        return lib().nanorq_generate_symbols(self.h, sbn, io.ptr)

    def encode(self, esi, sbn, io, out=None):
        T = self.symbol_size()
        if out is None:
            out = np.empty(T, dtype=np.uint8)
        n = lib().nanorq_encode(self.h, out.ctypes.data, esi, sbn, io.ptr)
        return out if n == T else None

    def encode_range(self, sbn, esi0, n, io, out=None):
        """nanorq_encode_range: symbols esi0..esi0+n-1 as an (n, T) array (or into `out`, whose rows
        may be wider than T)."""
        T = self.symbol_size()
        if out is None:
            out = np.empty((n, T), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.shape[0] >= n and out.strides[1] == 1 and out.shape[1] >= T
        got = lib().nanorq_encode_range(self.h, sbn, esi0, n, out.ctypes.data, out.strides[0], io.ptr)
        return out[:n, :T] if got == n else None

    def set_devices(self, n):
        return lib().nanorq_set_devices(self.h, n)


class Decoder(_Codec):
    def __init__(self, common, specific):
        super().__init__(lib().nanorq_decoder_new(common, specific))

    def set_max_esi(self, v):
        return lib().nanorq_set_max_esi(self.h, v)

    def add_symbol(self, data, tag, io):
        return lib().nanorq_decoder_add_symbol(self.h, data.ctypes.data, tag, io.ptr)

    def add_symbols(self, tags, data, io):
        """nanorq_decoder_add_symbols: data is an (n, >=T) uint8 array (rows may be strided);
        -> (number added or -1, per-symbol status list). `data` must stay alive until the blocks
        it feeds are complete."""
        tags = np.ascontiguousarray(tags, dtype=np.uint32)
        assert data.dtype == np.uint8 and data.shape[0] >= len(tags) and data.strides[1] == 1
        status = (C.c_int * max(1, len(tags)))()
        rc = lib().nanorq_decoder_add_symbols(self.h, tags.ctypes.data_as(u32p), data.ctypes.data, data.strides[0],
                                              len(tags), status, io.ptr)
        return rc, list(status[:len(tags)])

    def set_devices(self, n):
        return lib().nanorq_set_devices(self.h, n)

    def num_missing(self, sbn):
        return lib().nanorq_num_missing(self.h, sbn)

    def num_repair(self, sbn):
        return lib().nanorq_num_repair(self.h, sbn)

    def repair_block(self, io, sbn):
        return lib().nanorq_repair_block(self.h, io.ptr, sbn)

    def repair_blocks(self, io, sbns):
        """nanorq_repair_blocks: -> list of per-block results (one solve launch per device)."""
        arr = np.ascontiguousarray(sbns, dtype=np.uint8)
        ok = (C.c_bool * max(1, len(arr)))()
        lib().nanorq_repair_blocks(self.h, io.ptr, arr.ctypes.data_as(u8p), len(arr), ok)
        return list(ok[:len(arr)])


def tag(sbn, esi):
    return lib().nanorq_tag(sbn, esi)


class Solver:
    """rqb_solver: one source block resident on the GPU."""
    h = None

    def __init__(self, K, T, max_in=None, max_out=1, K_params=None, device=-1, flavour="auto"):
        self.K, self.T = K, T
        self.max_in = int(max_in or K)
        self.max_out = int(max_out)
        h = vp()
        _check(lib().rqb_solver_create_on(C.byref(h), device, K, K_params or K, T, self.max_in, self.max_out),
               "rqb_solver_create")
        self.h = h
        lib().rqb_solver_set_flavour(h, 1 if flavour == "hbm" else 0)
        lib().rqb_solver_set_timing(h, 1)  # tests and tools read last_kernel_ms()
        self.pitch = lib().rqb_solver_pitch(h)
        buf = (C.c_uint8 * (self.max_in * self.pitch)).from_address(lib().rqb_solver_staging(h))
        self.staging = np.frombuffer(buf, dtype=np.uint8).reshape(self.max_in, self.pitch)

    def close(self):
        if self.h:
            self.staging = None
            lib().rqb_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def upload(self, first, n):
        _check(lib().rqb_solver_upload(self.h, first, n), "rqb_solver_upload")

    def plan(self, req):
        rc = lib().rqb_solver_plan(self.h, C.byref(req.c))
        if rc not in (0, RQB_NEED_MORE):
            _check(rc, "rqb_solver_plan")
        return rc

    def plan_encode(self, want_c=True, n_repair=0):
        _check(lib().rqb_solver_plan_encode(self.h, 1 if want_c else 0, n_repair), "rqb_solver_plan_encode")

    def run(self):
        _check(lib().rqb_solver_run(self.h), "rqb_solver_run")

    def emit(self, isi):
        isi = _u32(isi)
        _check(lib().rqb_solver_emit(self.h, isi.ctypes.data_as(u32p), len(isi)), "rqb_solver_emit")

    def sync(self):
        _check(lib().rqb_solver_sync(self.h), "rqb_solver_sync")

    def fetch_syms(self, n, first=0):
        out = np.empty((n, self.T), dtype=np.uint8)
        _check(lib().rqb_solver_fetch_syms(self.h, first, n, out.ctypes.data, self.T), "rqb_solver_fetch_syms")
        return out

    def fetch_syms_mirror(self, n, first=0):
        _check(lib().rqb_solver_fetch_syms(self.h, first, n, None, 0), "rqb_solver_fetch_syms")

    def fetch_c(self, n=None, first=0):
        n = n if n is not None else block_params(self.K).L
        out = np.empty((n, self.T), dtype=np.uint8)
        _check(lib().rqb_solver_fetch_c(self.h, first, n, out.ctypes.data, self.T), "rqb_solver_fetch_c")
        return out

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(lib().rqb_solver_last_kernel_ms(self.h, C.byref(ms)), "rqb_solver_last_kernel_ms")
        return ms.value

    def mark(self, end=False):
        _check(lib().rqb_solver_mark(self.h, 1 if end else 0), "rqb_solver_mark")

    def marked_ms(self):
        ms = C.c_float()
        _check(lib().rqb_solver_marked_ms(self.h, C.byref(ms)), "rqb_solver_marked_ms")
        return ms.value

    def stats(self):
        st = SolverStats()
        _check(lib().rqb_solver_get_stats(self.h, C.byref(st)), "rqb_solver_get_stats")
        return st.as_dict()

    @staticmethod
    def plan_batch(solvers, requests, nthreads):
        """rqb_solver_plan_batch: plan requests[k] on solvers[k], up to nthreads host threads in C.
        -> list of per-block return codes (0 ok, 1 = needs more symbols)."""
        n = len(solvers)
        arr = (vp * n)(*[s.h for s in solvers])
        reqs = (_SolveRequest * n)(*[r.c for r in requests])
        rc = (C.c_int * n)()
        lib().rqb_solver_plan_batch(arr, reqs, n, int(nthreads), rc)
        return list(rc)

    @staticmethod
    def run_batch(solvers, owner=None):
        arr = (vp * len(solvers))(*[s.h for s in solvers])
        if owner is None:
            _check(lib().rqb_solver_run_batch(arr, len(solvers)), "rqb_solver_run_batch")
        else:
            _check(lib().rqb_solver_run_batch_on(arr, len(solvers), owner.h), "rqb_solver_run_batch_on")


class Matrix:
    """rqb_matrix: rows x T bytes in HBM, target of the batched row ops."""
    h = None

    def __init__(self, rows, T):
        self.rows, self.T = rows, T
        h = vp()
        _check(lib().rqb_matrix_create(C.byref(h), rows, T), "rqb_matrix_create")
        self.h = h
        self.pitch = lib().rqb_matrix_pitch(h)

    def close(self):
        if self.h:
            lib().rqb_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def upload(self, arr, first=0):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        _check(lib().rqb_matrix_upload(self.h, first, arr.shape[0], arr.ctypes.data, arr.strides[0]),
               "rqb_matrix_upload")

    def download(self, first=0, n=None):
        n = self.rows - first if n is None else n
        out = np.empty((n, self.T), dtype=np.uint8)
        _check(lib().rqb_matrix_download(self.h, first, n, out.ctypes.data, self.T), "rqb_matrix_download")
        return out

    def fill_random(self, seed=1):
        _check(lib().rqb_matrix_fill_random(self.h, seed), "rqb_matrix_fill_random")

    @staticmethod
    def make_ops(beta, i, j):
        ops = np.zeros(len(beta), dtype=OP_DTYPE)
        ops["beta"], ops["i"], ops["j"] = beta, i, j
        return ops

    def apply(self, ops):
        assert ops.dtype == OP_DTYPE
        _check(lib().rqb_rowops_apply(self.h, ops.ctypes.data, len(ops)), "rqb_rowops_apply")

    def upload_ops(self, ops):
        assert ops.dtype == OP_DTYPE
        h = vp()
        _check(lib().rqb_ops_upload(C.byref(h), ops.ctypes.data, len(ops)), "rqb_ops_upload")
        return h

    def apply_dev(self, oplist, repeats=1):
        ms = C.c_float()
        _check(lib().rqb_rowops_apply_dev(self.h, oplist, repeats, C.byref(ms)), "rqb_rowops_apply_dev")
        return ms.value

    def schedule_replay(self, ops, mark0, mark1, di, c, stepwise=False):
        assert ops.dtype == OP_DTYPE
        di = np.ascontiguousarray(di, dtype=np.int32)
        c = np.ascontiguousarray(c, dtype=np.int32)
        ms = C.c_float()
        fn = lib().rqb_schedule_replay_stepwise if stepwise else lib().rqb_schedule_replay
        _check(fn(self.h, ops.ctypes.data, len(ops), mark0, mark1,
                                         di.ctypes.data_as(C.POINTER(C.c_int)), len(di),
                                         c.ctypes.data_as(C.POINTER(C.c_int)), len(c), C.byref(ms)),
               "rqb_schedule_replay")
        return ms.value
