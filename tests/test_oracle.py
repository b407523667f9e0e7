"""CPU: pin the oracle (oracle/rq_oracle.c) against the committed golden fixtures
(generated from the unmodified reference by tests/golden/make_golden.py) and,
when the compiled reference is present, against oracle/_ref directly."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle_lib import (Op, fnv1a64, have_ref, kat_payload, oracle, orc_decode, orc_encode, orc_lt,
                        orc_params, ptr, ref, u32p)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat.json")))
SMALL = np.load(os.path.join(GOLD, "small.npz"))


def test_gf256_tables_against_shift_add():
    O = oracle()

    def slow(a, b):
        r = 0
        while b:
            if b & 1:
                r ^= a
            a = ((a << 1) ^ (0x11D if a & 0x80 else 0)) & 0xFF
            b >>= 1
        return r
    for a in range(256):
        for b in (0, 1, 2, 3, 29, 127, 128, 200, 255):
            assert O.orc_gf_mul(a, b) == slow(a, b)
        if a:
            assert O.orc_gf_mul(a, O.orc_gf_inv(a)) == 1


@pytest.mark.parametrize("K,T", [(10, 64), (1024, 1280), (4096, 1280)])
def test_oracle_reproduces_reference_kat(K, T):
    """SURVEY 8(c) hashes: intermediate symbols, repair symbols, op counts."""
    g = KAT["%d,%d" % (K, T)]
    p = orc_params(K)
    src = kat_payload(K * T)
    assert "%016x" % fnv1a64(src) == g["fnv_source"]
    Cm, nops, napp = orc_encode(K, T, src)
    assert nops == g["nops"]
    assert napp == g["nops"] + 2 * (g["marks"][0] + 1)
    assert "%016x" % fnv1a64(Cm) == g["fnv_intermediate"]
    rep = np.concatenate([orc_lt(K, T, Cm, e + p.Kprime - K) for e in range(K, K + 16)])
    assert "%016x" % fnv1a64(rep) == g["fnv_repair16"]


@pytest.mark.parametrize("key", ["K10_T64", "K26_T16", "K101_T24", "K257_T8"])
def test_oracle_reproduces_small_golden_vectors(key):
    K, T = (int(x[1:]) for x in key.split("_"))
    src = SMALL[key + "_src"]
    Cm, _, _ = orc_encode(K, T, src)
    assert np.array_equal(Cm, SMALL[key + "_C"])
    p = orc_params(K)
    esis, syms = SMALL[key + "_esis"], SMALL[key + "_syms"]
    for e, s in zip(esis, syms):  # every symbol the reference emitted
        want = src.reshape(K, T)[e] if e < K else orc_lt(K, T, Cm, int(e) + p.Kprime - K)
        assert np.array_equal(s, want)
    rc, out, _, _, _ = orc_decode(K, T, esis, syms)
    assert rc == int(SMALL[key + "_rc"][0])
    if rc == 0:
        assert np.array_equal(out, SMALL[key + "_out"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("K", [10, 12, 55, 101, 500, 1024, 2000])
def test_oracle_op_list_identical_to_reference(K):
    R, O = ref(), oracle()
    T = 16
    p = orc_params(K)
    rp = (C.c_int * 10)()
    R.ref_params(K, C.byref(rp))
    assert list(rp) == [p.Kprime, p.S, p.H, p.W, p.L, p.P, p.P1, p.U, p.B, p.J]
    rng = np.random.default_rng(K)
    src = rng.integers(0, 256, (K, T), dtype=np.uint8)
    D = np.zeros((p.L, T), np.uint8)
    D[p.S + p.H:p.S + p.H + K] = src
    isi = np.arange(p.Kprime, dtype=np.uint32)
    Cr = np.zeros((p.L, T), np.uint8)
    info = (C.c_long * 5)()
    ops = np.zeros((400000, 3), np.uint32)
    assert R.ref_solve(K, T, 0, ptr(isi, u32p), ptr(D), ptr(Cr), C.byref(info), ptr(ops, u32p), len(ops)) == 0
    st = C.c_int()
    S = O.orc_invert(C.byref(p), 0, ptr(isi, u32p), C.byref(st))
    s = S.contents
    assert [s.nops, s.marks[0], s.marks[1], s.i, s.u] == list(info)
    mine = np.array([(s.ops[k].beta, s.ops[k].i, s.ops[k].j) for k in range(s.nops)], dtype=np.uint32)
    assert np.array_equal(mine, ops[:s.nops])
    O.orc_sched_free(S)
    Cm, _, _ = orc_encode(K, T, src)
    assert np.array_equal(Cm, Cr)
    for x in (0, 1, K - 1, K, K + 7, 3 * K):
        idx_r, idx_o = (C.c_uint32 * 40)(), (C.c_uint32 * 40)()
        n = R.ref_lt_indices(K, x, idx_r)
        assert n == O.orc_lt_indices(C.byref(p), x, idx_o) and list(idx_r[:n]) == list(idx_o[:n])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("K,T,loss,oh,trials", [(10, 8, 0.4, 0, 150), (26, 8, 0.5, 0, 60), (100, 16, 0.3, 12, 5),
                                                  (1024, 32, 0.05, 2, 2)])
def test_oracle_decode_verdict_and_bytes_match_reference_api(K, T, loss, oh, trials):
    R = ref()
    for seed in range(trials):
        rng = np.random.default_rng(1000 * K + seed)
        src = rng.integers(0, 256, K * T, dtype=np.uint8)
        drop = rng.random(K) < loss
        esis = np.concatenate([np.nonzero(~drop)[0], np.arange(K, K + int(drop.sum()) + oh)]).astype(np.uint32)
        rng.shuffle(esis)
        syms = np.zeros((len(esis), T), np.uint8)
        oti = (C.c_uint64 * 2)()
        assert R.ref_encode_api(K, T, ptr(src), ptr(esis, u32p), len(esis), ptr(syms), C.byref(oti), None, None, 0) == 0
        out = np.zeros(K * T, np.uint8)
        rc_ref = R.ref_decode_api(C.byref(oti), T, ptr(esis, u32p), ptr(syms), len(esis), ptr(out), K * T, None, None)
        rc_o, out_o, _, _, _ = orc_decode(K, T, esis, syms)
        assert (rc_ref == 0) == (rc_o == 0)
        if rc_ref == 0:
            assert np.array_equal(out, src) and np.array_equal(out_o, src)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_row_kernels_match_oblas_avx_for_all_multipliers():
    """oaxpy / oscal of the AVX build vs the scalar restatement, all 256 u."""
    R, O = ref(), oracle()
    k = 96  # ALIGNED_COLS(96) = 96 with OCTMAT_ALIGN 32
    rng = np.random.default_rng(7)
    for u in range(256):
        a = rng.integers(0, 256, (4, k), dtype=np.uint8)
        b = a.copy()
        R.oaxpy(ptr(a), ptr(a), 1, 2, k, u)
        O.orc_row_axpy(ptr(b[1]), ptr(b[2]), k, u)
        R.oscal(ptr(a), 3, k, u)
        O.orc_row_scal(ptr(b[3]), k, u)
        assert np.array_equal(a, b), u
