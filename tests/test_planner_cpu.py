"""CPU: the host half of the product -- planner, program format, API glue --
checked against the oracle through the CPU interpreter of the device program
(oracle/plan_interp.c).  No GPU compute happens here."""
import ctypes as C
import os

import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import api
from oracle_lib import have_ref, interp_run, kat_payload, orc_decode, orc_encode, orc_lt, orc_params, ref


def test_library_exports_every_declared_symbol():
    L = nb.lib()
    missing = [s for s in api.EXPORTED_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    # every prototype in include/*.h is covered by the list
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    import re
    declared = set()
    for fn in os.listdir(inc):
        txt = open(os.path.join(inc, fn)).read()
        declared |= set(re.findall(r"\b((?:nanorq|rqb|ioctx)_[a-z0-9_]+)\s*\(", txt))
    declared -= {"rqb_solver_plan_encode_"}
    assert declared <= set(api.EXPORTED_SYMBOLS), declared - set(api.EXPORTED_SYMBOLS)


@pytest.mark.skipif(nb.device_count() > 0, reason="only meaningful without a GPU")
def test_product_fails_loudly_without_a_gpu():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.Solver(10, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.Matrix(4, 64)
    enc = nb.Encoder(640, 64, 10, 0, 8)
    io = nb.MemIO(np.zeros(640, np.uint8))
    assert enc.generate_symbols(0, io) is False
    assert enc.encode(11, 0, io) is None


@pytest.mark.parametrize("K", [1, 9, 10, 11, 4096, 4112, 56403])
def test_block_params_match_oracle(K):
    p, q = nb.block_params(K), orc_params(K)
    for f, _ in api.BlockParams._fields_:
        assert getattr(p, f) == getattr(q, f), f
    for x in (0, K, 5 * K + 3, (1 << 24) - 1):
        import oracle_lib as ol
        out = (C.c_uint32 * 40)()
        n = ol.oracle().orc_lt_indices(C.byref(q), x, out)
        assert nb.lt_row_indices(K, x) == list(out[:n])


@pytest.mark.parametrize("smem", [False, True], ids=["hbm", "smem"])
@pytest.mark.parametrize("K,T", [(10, 64), (26, 16), (101, 24), (257, 8), (1024, 16), (4096, 8)])
def test_encode_plan_on_interpreter_equals_oracle(K, T, smem):
    p = orc_params(K)
    src = kat_payload(K * T).reshape(K, T)
    Cm, _, _ = orc_encode(K, T, src)
    out_isi = np.arange(K, K + 24, dtype=np.uint32) + (p.Kprime - K)
    req = nb.SolveRequest.for_encoder(K, True, out_isi)
    rc, blob = nb.plan_blob(K, req, smem=smem)
    assert rc == 0
    assert blob["smem"] == int(smem)  # every one of these blocks fits the shared-memory slots
    st = blob["stats"]
    assert st["i"] + st["u"] == p.L and st["rho"] + st["nfree"] == st["u"]
    rc, cout, sout = interp_run(blob, src, T, p.L, len(out_isi))
    assert rc == 0  # 10 would mean an intra-level hazard in the program
    assert np.array_equal(cout, Cm)
    assert np.array_equal(sout, np.stack([orc_lt(K, T, Cm, int(x)) for x in out_isi]))


@pytest.mark.parametrize("smem", [False, True], ids=["hbm", "smem"])
@pytest.mark.parametrize("K,T,loss,oh,trials", [(10, 8, 0.4, 0, 200), (26, 8, 0.5, 0, 80), (100, 16, 0.5, 0, 4),
                                                  (100, 16, 0.2, 12, 4), (257, 8, 0.3, 15, 3),
                                                  (1000, 16, 0.9, 1, 2), (1024, 32, 0.05, 2, 2), (4096, 8, 0.1, 0, 1)])
def test_decode_plan_on_interpreter_equals_oracle_including_verdict(K, T, loss, oh, trials, smem):
    p = orc_params(K)
    singular = 0
    for seed in range(trials):
        rng = np.random.default_rng(77 * K + seed)
        src = rng.integers(0, 256, (K, T), dtype=np.uint8)
        Cm, _, _ = orc_encode(K, T, src)
        drop = rng.random(K) < loss
        esis = np.concatenate([np.nonzero(~drop)[0], np.arange(K, K + int(drop.sum()) + oh)]).astype(np.uint32)
        rng.shuffle(esis)
        syms = np.stack([src[e] if e < K else orc_lt(K, T, Cm, int(e) + p.Kprime - K) for e in esis])
        rc_o, out_o, C_o, _, _ = orc_decode(K, T, esis, syms, want_C=True)
        req, missing = nb.SolveRequest.for_decoder(K, esis)
        if not missing:
            continue
        rc_p, blob = nb.plan_blob(K, req, smem=smem)
        assert (rc_o == 0) == (rc_p == 0), (K, seed, rc_o, rc_p)
        if rc_o != 0:
            singular += 1
            continue
        rc, cout, sout = interp_run(blob, syms, T, p.L, len(missing))
        assert rc == 0
        assert np.array_equal(cout, C_o)
        assert np.array_equal(sout, src[missing])
    print("singular cases:", singular)


def test_plan_reports_rank_deficiency_like_the_reference():
    """K=10 with the same symbol fed twice cannot be solved: one equation short."""
    K, T = 10, 8
    esis = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8], dtype=np.uint32)  # 9 of 10, no repair
    req, missing = nb.SolveRequest.for_decoder(K, esis)
    assert req is None and missing == [9]


@pytest.mark.parametrize("args", [(640, 64, 10, 0, 8), (5242880, 1280, 4096, 0, 8), (1000003, 1280, 0, 0, 4),
                                  (77777, 100, 0, 5, 8), (28878336, 512, 56403, 0, 8), (123, 17, 3, 0, 3),
                                  (41943040, 1280, 4096, 0, 8), (999, 7, 0, 0, 1)])
def test_oti_and_partitioning_match_reference(args):
    enc = nb.Encoder(*args)
    common, scheme = enc.oti_common(), enc.oti_scheme_specific()
    dec = nb.Decoder(common, scheme)
    assert dec.blocks() == enc.blocks() and dec.symbol_size() == enc.symbol_size()
    assert dec.transfer_length() == enc.transfer_length() == args[0]
    sizes = [enc.block_symbols(b) for b in range(enc.blocks())]
    assert sizes == [dec.block_symbols(b) for b in range(dec.blocks())]
    T = enc.symbol_size()
    assert sum(sizes) == -(-args[0] // T)
    if have_ref():
        R = ref()
        R.nanorq_encoder_new_ex.restype = C.c_void_p
        R.nanorq_encoder_new_ex.argtypes = [C.c_size_t, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint8]
        for f, rt in (("nanorq_oti_common", C.c_uint64), ("nanorq_oti_scheme_specific", C.c_uint32),
                      ("nanorq_blocks", C.c_size_t), ("nanorq_symbol_size", C.c_size_t)):
            getattr(R, f).restype = rt
            getattr(R, f).argtypes = [C.c_void_p]
        R.nanorq_block_symbols.restype = C.c_size_t
        R.nanorq_block_symbols.argtypes = [C.c_void_p, C.c_uint8]
        R.nanorq_free.argtypes = [C.c_void_p]
        h = R.nanorq_encoder_new_ex(*args)
        assert h
        assert R.nanorq_oti_common(h) == common and R.nanorq_oti_scheme_specific(h) == scheme
        assert R.nanorq_blocks(h) == enc.blocks() and R.nanorq_symbol_size(h) == T
        assert [R.nanorq_block_symbols(h, b) for b in range(enc.blocks())] == sizes
        R.nanorq_free(h)


def test_constructor_argument_errors():
    with pytest.raises(ValueError):
        nb.Encoder(0, 64, 10, 0, 8)
    with pytest.raises(ValueError):
        nb.Encoder(946270874881, 1280, 0, 0, 8)
    with pytest.raises(ValueError):
        nb.Encoder(10 ** 9, 8, 1, 0, 8)  # more than 256 blocks
    with pytest.raises(ValueError):
        nb.Decoder((640 << 24) | 63, 0)  # Al = 0
    assert api.tag(3, 0x01000005) == (3 << 24) | 5


@pytest.mark.parametrize("K,T", [(10, 16), (26, 8), (257, 8), (1024, 8)])
def test_reference_schedule_as_one_program_equals_oracle(K, T):
    """rqb_schedule_replay's program (reference-format sched_op list -> one levelled gather
    program, accumulations merged) run on the interpreter: the matrix becomes the
    intermediate symbols, and the merged program has far fewer levels than ops."""
    import ctypes as C
    from nanorq_b200 import api
    from oracle_lib import oracle, ptr, u32p
    O = oracle()
    p = orc_params(K)
    src = kat_payload(K * T).reshape(K, T)
    isi = np.arange(p.Kprime, dtype=np.uint32)
    st = C.c_int()
    S = O.orc_invert(C.byref(p), 0, ptr(isi, u32p), C.byref(st))
    s = S.contents
    ops = np.zeros(s.nops, dtype=api.OP_DTYPE)
    C.memmove(ops.ctypes.data, s.ops, s.nops * 12)
    di = np.ctypeslib.as_array(s.di, (s.rows,)).copy()
    c = np.ctypeslib.as_array(s.c, (s.cols,)).copy()
    rc, blob = nb.schedule_plan_blob(p.L, ops, s.marks[0], s.marks[1], di, c)
    n_applied = s.nops + 2 * (s.marks[0] + 1)
    O.orc_sched_free(S)
    assert rc == 0
    D = np.zeros((p.L, T), np.uint8)
    D[p.S + p.H:p.S + p.H + K] = src
    rc, cout, _ = interp_run(blob, D, T, p.L, 0, in_writable=True)
    assert rc == 0
    want, _, _ = orc_encode(K, T, src)
    assert np.array_equal(cout, want)
    assert blob["stats"]["n_tasks"] <= n_applied + p.L and blob["stats"]["n_levels"] < n_applied
    print("K=%d: %d applied ops -> %d tasks in %d levels" % (K, n_applied, blob["stats"]["n_tasks"], blob["stats"]["n_levels"]))


@pytest.mark.parametrize("seed", range(6))
def test_random_op_list_as_one_program_equals_sequential_oracle(seed):
    """Arbitrary reference-format op lists (axpy with beta 1 and beta > 1, oscal including the
    no-op multipliers 0 and 1, a row added to itself, long accumulation runs, read-after-write
    and write-after-read chains) turned into one program must equal the ops applied one by
    one by the oracle's row kernels; plus a random final row permutation."""
    import ctypes as C
    from nanorq_b200 import api
    from oracle_lib import Op, oracle
    rng = np.random.default_rng(100 + seed)
    nr, T, n = int(rng.integers(3, 40)), 24, int(rng.integers(1, 900))
    D = rng.integers(0, 256, (nr, T), dtype=np.uint8)
    ops = np.zeros(n, dtype=api.OP_DTYPE)
    hot = rng.integers(0, nr, 3)  # a few destinations that accumulate a lot
    for k in range(n):
        kind = rng.random()
        i = int(hot[rng.integers(0, 3)]) if rng.random() < 0.4 else int(rng.integers(0, nr))
        if kind < 0.08:
            ops[k] = (0, i, int(rng.integers(0, 256)))  # oscal, multiplier in j (0 and 1 are no-ops)
        elif kind < 0.10:
            ops[k] = (int(rng.integers(1, 256)), i, i)  # row added to itself
        else:
            ops[k] = (1 if rng.random() < 0.6 else int(rng.integers(2, 256)), i, int(rng.integers(0, nr)))
    perm = rng.permutation(nr).astype(np.int32)
    ident = np.arange(nr, dtype=np.int32)
    # oracle: one op at a time, then the cycle-walk permutation of precode_matrix_permute
    want = D.copy()
    arr = (Op * n)()
    for k in range(n):
        arr[k].beta, arr[k].i, arr[k].j = int(ops["beta"][k]), int(ops["i"][k]), int(ops["j"][k])
    oracle().orc_apply_ops(want.ctypes.data_as(C.POINTER(C.c_uint8)), T, T, arr, n)
    P = perm.copy()
    for i in range(nr):  # lib/precode.c:3-13
        at = i
        while P[at] >= 0:
            tmp = want[i].copy(); want[i] = want[P[at]]; want[P[at]] = tmp
            nxt = P[at]; P[at] = -1; at = nxt
    rc, blob = nb.schedule_plan_blob(nr, ops, -1, 0, perm, ident)
    assert rc == 0
    rc, cout, _ = interp_run(blob, D, T, nr, 0, in_writable=True)
    assert rc == 0
    assert np.array_equal(cout, want)


@pytest.mark.parametrize("smem", [False, True], ids=["hbm", "smem"])
def test_out_row_places_recovered_symbols_at_their_esi(smem):
    """rqb_solve_request.out_row: a decoder recovers straight into the rows of the block image."""
    K, T = 200, 16
    p = orc_params(K)
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (K, T), dtype=np.uint8)
    Cm, _, _ = orc_encode(K, T, src)
    drop = rng.random(K) < 0.2
    esis = np.concatenate([np.nonzero(~drop)[0], np.arange(K, K + int(drop.sum()) + 2)]).astype(np.uint32)
    syms = np.stack([src[e] if e < K else orc_lt(K, T, Cm, int(e) + p.Kprime - K) for e in esis])
    req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
    req2 = nb.SolveRequest(req.isi, req.in_row, req.c.overhead, False, missing, out_row=missing)
    # the blob sizes the emitted-symbol space to n_out rows: rows beyond it are refused
    assert max(missing) >= len(missing)
    assert nb.plan_blob(K, req2, smem=smem)[0] != 0
    bad = nb.SolveRequest(req.isi, req.in_row, req.c.overhead, False, missing, out_row=[10 ** 6] * len(missing))
    assert nb.plan_blob(K, bad, smem=smem)[0] != 0
    # ... and a permutation inside the space is honoured
    perm = np.random.default_rng(1).permutation(len(missing)).astype(np.uint32)
    req3 = nb.SolveRequest(req.isi, req.in_row, req.c.overhead, False, missing, out_row=perm)
    rc, blob = nb.plan_blob(K, req3, smem=smem)
    assert rc == 0
    rc, _, sout = interp_run(blob, syms, T, 0, len(missing))
    assert rc == 0
    assert np.array_equal(sout[perm], src[missing])
