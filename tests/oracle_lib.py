"""ctypes bindings for the TEST-SIDE checkers: oracle/liboracle.so (this repo's C
restatement) and oracle/_ref/libnanorq_ref.so (the unmodified reference compiled
from /root/reference by oracle/Makefile).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnanorq_ref.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Kprime", "S", "H", "W", "L", "P", "P1", "U", "B", "J")]


class Op(C.Structure):
    _fields_ = [("beta", C.c_uint8), ("i", C.c_uint32), ("j", C.c_uint32)]


class Sched(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int),
                ("c", C.POINTER(C.c_int)), ("ci", C.POINTER(C.c_int)),
                ("d", C.POINTER(C.c_int)), ("di", C.POINTER(C.c_int)),
                ("ops", C.POINTER(Op)), ("nops", C.c_size_t), ("cap", C.c_size_t),
                ("i", C.c_int), ("u", C.c_int), ("marks", C.c_long * 2), ("promotions", C.c_int)]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


_oracle = None
_ref = None


def ptr(a, t=u8p):
    return a.ctypes.data_as(t)


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_params_init.argtypes = [C.c_int, C.POINTER(Params)]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [u8p, C.c_size_t]
        L.orc_rand.restype = C.c_uint32
        L.orc_rand.argtypes = [C.c_uint32] * 3
        L.orc_lt_indices.argtypes = [C.POINTER(Params), C.c_uint32, u32p]
        L.orc_gf_mul.restype = C.c_uint8
        L.orc_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
        L.orc_gf_inv.restype = C.c_uint8
        L.orc_gf_inv.argtypes = [C.c_uint8]
        L.orc_row_axpy.argtypes = [u8p, u8p, C.c_size_t, C.c_uint8]
        L.orc_row_scal.argtypes = [u8p, C.c_size_t, C.c_uint8]
        L.orc_apply_ops.argtypes = [u8p, C.c_size_t, C.c_size_t, C.POINTER(Op), C.c_size_t]
        L.orc_invert.restype = C.POINTER(Sched)
        L.orc_invert.argtypes = [C.POINTER(Params), C.c_int, u32p, C.POINTER(C.c_int)]
        L.orc_sched_free.argtypes = [C.POINTER(Sched)]
        L.orc_applied_ops.restype = C.c_size_t
        L.orc_applied_ops.argtypes = [C.POINTER(Sched), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.orc_intermediate.argtypes = [C.POINTER(Sched), u8p, C.c_size_t, C.c_size_t]
        L.orc_lt_row.argtypes = [C.POINTER(Params), u8p, C.c_size_t, C.c_uint32, u8p, C.c_size_t]
        L.orc_encode_block.argtypes = [C.c_int, C.c_size_t, u8p, u8p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.orc_decode_block.argtypes = [C.c_int, C.c_size_t, u32p, u8p, C.c_size_t, u8p, u8p,
                                       C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_params.argtypes = [C.c_int, C.POINTER(C.c_int * 10)]
        L.ref_lt_indices.argtypes = [C.c_int, C.c_uint32, u32p]
        L.ref_rand.restype = C.c_uint32
        L.ref_rand.argtypes = [C.c_uint32] * 3
        L.ref_solve.argtypes = [C.c_int, C.c_size_t, C.c_int, u32p, u8p, u8p,
                                C.POINTER(C.c_long * 5), u32p, C.c_size_t]
        L.ref_lt_row.argtypes = [C.c_int, C.c_size_t, u8p, C.c_uint32, u8p]
        L.ref_encode_api.argtypes = [C.c_size_t, C.c_size_t, u8p, u32p, C.c_size_t, u8p,
                                     C.POINTER(C.c_uint64 * 2), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_int]
        L.ref_decode_api.argtypes = [C.POINTER(C.c_uint64 * 2), C.c_size_t, u32p, u8p, C.c_size_t,
                                     u8p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in ("oaxpy",):
            getattr(L, f).argtypes = [u8p, u8p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint8]
        L.oaddrow.argtypes = [u8p, u8p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.oscal.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint8]
        _ref = L
    return _ref


# ---------------------------------------------------------------- helpers
def kat_payload(n):
    """in[i] = (uint8)(((uint32)i * 2654435761) >> 24)"""
    i = np.arange(n, dtype=np.uint64)
    return (((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(24)).astype(np.uint8)


def fnv1a64(a):
    """FNV-1a 64 over a byte array (the hash of the KAT fixtures)."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return int(oracle().orc_fnv1a64(ptr(a), a.size))


INTERP_SO = os.path.join(ROOT, "oracle", "libplan_interp.so")
_interp = None


def interp_run(blob, inp, T, n_c, n_out, in_writable=False):
    """Run a plan blob (nanorq_b200.plan_blob) on the CPU interpreter.
    -> (rc, C[n_c,T], syms[n_out,T]); rc 10 = hazard inside a level, 12 = misaligned,
    13 = more than RQB_MAX_SRCS sources in a task, 14 = writes the input space / ZERO row,
    15 = an XOR source list is not padded with the ZERO row."""
    global _interp
    if _interp is None:
        _interp = C.CDLL(INTERP_SO)
        sz = C.c_size_t
        _interp.rqb_interp_run2.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32, u8p, u8p, sz, sz, sz,
                                            u8p, sz, sz, u8p, sz, sz, C.c_int, C.c_int, C.c_uint32]
    inp = np.ascontiguousarray(inp, dtype=np.uint8)
    row0 = np.asarray(blob["row0"], dtype=np.uint32)
    in_rows = min(inp.shape[0], int(row0[1]))  # rows past the plan's input space are never referenced
    cout = np.full((max(n_c, 1), T), 0x5A, np.uint8)
    sout = np.full((max(n_out, 1), T), 0x5A, np.uint8)
    # both flavours of the program (rqb_program.h): blob["smem"] selects the shared-memory one
    rc = _interp.rqb_interp_run2(ptr(row0, u32p), blob["zero_row"], blob["n_rows"], blob["n_pages"],
                                 ptr(blob["pages"]), ptr(inp), in_rows, inp.strides[0], T, ptr(cout), n_c, T,
                                 ptr(sout), n_out, T, 1 if in_writable else 0, int(blob.get("smem", 0)),
                                 int(blob.get("n_slots", 0)))
    return rc, cout[:n_c], sout[:n_out]


def orc_params(K):
    p = Params()
    assert oracle().orc_params_init(K, C.byref(p)) == 0
    return p


def orc_encode(K, T, src):
    """-> (C[L,T], nops, n_applied)"""
    p = orc_params(K)
    Cm = np.zeros((p.L, T), dtype=np.uint8)
    nops, napp = C.c_size_t(), C.c_size_t()
    rc = oracle().orc_encode_block(K, T, ptr(np.ascontiguousarray(src)), ptr(Cm), T,
                                   C.byref(nops), C.byref(napp))
    assert rc == 0, rc
    return Cm, nops.value, napp.value


def orc_lt(K, T, Cm, isi):
    p = orc_params(K)
    out = np.zeros(T, dtype=np.uint8)
    oracle().orc_lt_row(C.byref(p), ptr(Cm), Cm.strides[0], isi, ptr(out), T)
    return out


def orc_decode(K, T, esis, syms, want_C=False):
    """-> (rc, out[K*T], C or None, nops, n_applied)"""
    p = orc_params(K)
    esis = np.ascontiguousarray(esis, dtype=np.uint32)
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    out = np.zeros(K * T, dtype=np.uint8)
    Cm = np.zeros((p.L, T), dtype=np.uint8) if want_C else None
    nops, napp = C.c_size_t(), C.c_size_t()
    rc = oracle().orc_decode_block(K, T, ptr(esis, u32p), ptr(syms), len(esis), ptr(out),
                                   ptr(Cm) if want_C else None, T, C.byref(nops), C.byref(napp))
    return rc, out, Cm, nops.value, napp.value
