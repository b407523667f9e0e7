"""Seeded random sessions against nanorq.h + nanorq_batch.h on the GPU: random object shapes (ragged last
symbol, several source blocks), random loss, duplicates, shuffled arrival, symbols produced by a mix of
nanorq_encode / nanorq_encode_range and delivered by a mix of nanorq_decoder_add_symbol /
nanorq_decoder_add_symbols (with holes), early repair attempts, pageable and page-locked output.
Every session must end with the decoded object equal to the payload, and every repair symbol equal to
the oracle's (the reference's encoder, oracle/_ref or the C restatement)."""
import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import api
from oracle_lib import orc_decode, orc_encode, orc_lt, orc_params

pytestmark = pytest.mark.gpu


def session(seed, big=False, files=None):
    rng = np.random.default_rng(seed)
    T = int(rng.choice([8, 16, 24, 64, 104, 256, 1280]))
    kind = rng.integers(0, 4)
    if big:  # blocks of several MB: the HBM flavour, block images, several repair windows
        T = int(rng.choice([512, 1000, 1280, 1496]))
        K = int(rng.integers(1500, 9000))
    elif kind == 0:
        K = int(rng.integers(1, 40))       # tiny blocks: no device context until it is needed
    elif kind == 1:
        K = int(rng.integers(40, 300))     # around the lazy / eager boundary (256 symbols)
    else:
        K = int(rng.integers(300, 1400))
    if K * T > (1 << 20) and not big:
        K = (1 << 20) // T
    nblocks = int(rng.integers(1, 4)) if not big else int(rng.integers(1, 3))
    F = nblocks * K * T - int(rng.integers(0, T))  # the object's last symbol is short most of the time
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    enc = nb.Encoder(F, T, K, 0, 8)
    Z = enc.blocks()
    spread = nb.device_count() > 1 and rng.random() < 0.5  # the blocks of the object on all GPUs of the box
    if spread:
        assert enc.set_devices(0) == nb.device_count()
    in_kind = rng.choice(["mem", "file", "mmap"]) if files else "mem"
    if in_kind == "mem":
        io_in = nb.MemIO(payload)
    else:  # encode.c style: the object is read through a file ioctx
        payload.tofile(files / ("in%d.bin" % seed))
        io_in = nb.FileIO(files / ("in%d.bin" % seed), 1, mmap=in_kind == "mmap")
    loss = float(rng.choice([0.0, 0.05, 0.2, 0.5]))
    packets = []  # (tag, row)
    extra = {}    # sbn -> next unused repair ESI
    for sbn in range(Z):
        Kb = enc.block_symbols(sbn)
        drop = rng.random(Kb) < loss
        n_rep = int(drop.sum()) + int(rng.integers(0, 4))
        if rng.random() < 0.5:  # one range call
            syms = enc.encode_range(sbn, 0, Kb + n_rep, io_in)
            assert syms is not None, (seed, sbn, Kb, n_rep, T, api.last_error())
        else:                   # per-symbol calls, repair symbols first
            syms = np.zeros((Kb + n_rep, T), np.uint8)
            for esi in list(range(Kb, Kb + n_rep)) + list(range(Kb)):
                s = enc.encode(esi, sbn, io_in)
                assert s is not None
                syms[esi] = s
        if seed % 5 == 0:  # repair symbols against the oracle
            blk = np.zeros(Kb * T, np.uint8)
            first = sum(enc.block_symbols(q) for q in range(sbn)) * T
            blk[:min(Kb * T, F - first)] = payload[first:first + Kb * T]
            p = orc_params(Kb)
            Co, _, _ = orc_encode(Kb, T, blk)
            for esi in range(Kb, Kb + min(n_rep, 6)):
                assert np.array_equal(syms[esi], orc_lt(Kb, T, Co, esi + p.Kprime - Kb)), (seed, sbn, esi)
        for esi in list(np.nonzero(~drop)[0]) + list(range(Kb, Kb + n_rep)):
            packets.append((api.tag(sbn, int(esi)), syms[esi].copy()))
        extra[sbn] = Kb + n_rep
        if rng.random() < 0.3:  # the sender drops the block's state; a later request for more repair symbols reloads it
            enc.encoder_cleanup(sbn)
    # duplicates and arrival order
    for _ in range(int(rng.integers(0, 6))):
        packets.append(packets[int(rng.integers(0, len(packets)))])
    if rng.random() < 0.6:
        order = rng.permutation(len(packets))
        packets = [packets[i] for i in order]
    pinned = rng.random() < 0.5
    out_kind = rng.choice(["mem", "file", "mmap"]) if files else "mem"
    keep = []
    if out_kind != "mem":  # decode.c style: the decoded object is written through a file ioctx
        pinned = False
        out = None
        io = nb.FileIO(files / ("out%d.bin" % seed), 0, mmap=out_kind == "mmap")
    elif pinned:
        ob = nb.PinnedBuffer(F)
        out = ob.arr
        out[:] = 0xEE
        keep.append(ob)
        io = nb.PinnedMemIO(out)
    else:
        out = np.full(F, 0xEE, np.uint8)
        io = nb.MemIO(out)
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    if spread:
        assert dec.set_devices(0) == nb.device_count()
    k = 0
    while k < len(packets):
        if rng.random() < 0.5:  # a batch call over a ring with holes
            n = int(rng.integers(1, 80))
            chunk = packets[k:k + n]
            tags, rows = [], []
            for t, r in chunk:
                if rng.random() < 0.1:
                    tags.append(0xFFFFFFFF)
                    rows.append(np.zeros(T, np.uint8))
                tags.append(t)
                rows.append(r)
            if pinned and rng.random() < 0.5:
                pb = nb.PinnedBuffer(len(rows) * T)
                pb.arr[:] = np.stack(rows).reshape(-1)
                data = pb.arr.reshape(len(rows), T)
                keep.append(pb)
            else:
                data = np.stack(rows)
            rc, st = dec.add_symbols(np.array(tags, np.uint32), data, io)
            assert rc >= 0, (seed, rc)
            k += len(chunk)
        else:
            t, r = packets[k]
            assert dec.add_symbol(r, int(t), io) != nb.SYM_ERR, seed
            k += 1
        if rng.random() < 0.05:  # an impatient receiver: may or may not be decodable yet, must not break anything
            dec.repair_block(io, int(rng.integers(0, Z)))
    if rng.random() < 0.5:
        api.lib().rqb_set_plan_threads(int(rng.choice([1, 3])))  # the blocks' programs built side by side or in turn
        oks = dec.repair_blocks(io, list(range(Z)))
        api.lib().rqb_set_plan_threads(1)
    else:
        oks = [dec.repair_block(io, sbn) for sbn in range(Z)]
    for sbn in range(Z):  # a singular matrix: two more repair symbols, as a receiver would ask for
        tries = 0
        while not oks[sbn] and tries < 6:
            for _ in range(2):
                esi = extra[sbn]
                extra[sbn] += 1
                s = enc.encode(esi, sbn, io_in)
                assert dec.add_symbol(s, api.tag(sbn, esi), io) != nb.SYM_ERR
            oks[sbn] = dec.repair_block(io, sbn)
            tries += 1
        assert oks[sbn], (seed, sbn)
        assert dec.num_missing(sbn) == 0
    dec.close()
    io.close()
    if out is None:
        out = np.fromfile(files / ("out%d.bin" % seed), dtype=np.uint8)
    assert len(out) == F and np.array_equal(out, payload), seed
    enc.close()
    io_in.close()
    for b in keep:
        b.close()


@pytest.mark.parametrize("chunk", range(8))
def test_random_sessions(chunk):
    for seed in range(chunk * 40, chunk * 40 + 40):
        session(1000 + seed)


def test_random_sessions_through_file_ioctx(tmp_path):
    for seed in range(40):
        session(3000 + seed, files=tmp_path)


def test_random_sessions_with_large_blocks():
    for seed in range(16):
        session(7000 + seed, big=True)


def test_more_encoder_shapes_than_the_program_cache_holds():
    """Encoder programs are cached per block size (48 entries, least recently used dropped): 70 sizes in a
    row, each checked against the oracle, then the first sizes again."""
    T = 16
    sizes = list(range(300, 370)) + list(range(300, 310))
    for K in sizes:
        rng = np.random.default_rng(K)
        payload = rng.integers(0, 256, K * T, dtype=np.uint8)
        enc = nb.Encoder(K * T, T, K, 0, 8)
        io = nb.MemIO(payload)
        got = enc.encode_range(0, K, 3, io)
        p = orc_params(K)
        Co, _, _ = orc_encode(K, T, payload)
        for k in range(3):
            assert np.array_equal(got[k], orc_lt(K, T, Co, p.Kprime + k)), (K, k)
        enc.close()
        io.close()


def test_random_sessions_on_four_threads_at_once():
    """The same sessions from four threads at a time (ctypes releases the GIL inside the library):
    mixed block shapes contend for the context list, the buffer pool and the program cache."""
    import threading
    errors = []

    def run(first):
        try:
            for seed in range(first, first + 25):
                session(5000 + seed)
        except BaseException as e:  # noqa: BLE001 -- reported in the main thread
            errors.append((first, repr(e)))

    threads = [threading.Thread(target=run, args=(100 * t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_random_sessions_with_a_tiny_cache():
    """A cache limit of 4 MiB: nearly every retired context and buffer is really freed and the next
    block allocates afresh -- the eviction paths run all the time, results stay right, and what is
    kept stays under the limit."""
    _, limit = nb.cache_stats()
    nb.release_cached()
    nb.set_cache_limit(4 << 20)
    try:
        for seed in range(60):
            session(9000 + seed)
            assert nb.cache_stats()[0] <= (4 << 20)
    finally:
        nb.set_cache_limit(limit)
        nb.release_cached()


@pytest.mark.parametrize("K,T,loss", [(10, 64, 0.0), (10, 64, 0.3), (1024, 1280, 0.05), (300, 104, 0.2)])
def test_the_call_pattern_of_the_reference_benchmark(K, T, loss):
    """benchmark.c:82-170 of the reference: the encoder loop calls nanorq_generate_symbols for every block and
    then nanorq_encoder_reset(rq, 0); the decoder loop adds ALL packets again every round (blocks that are
    already complete ignore theirs), repairs every block and resets block 0 only."""
    rng = np.random.default_rng(K + T)
    nblocks = 3
    F = nblocks * K * T
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    enc = nb.Encoder(F, T, K, 0, 8)
    io_in = nb.MemIO(payload)
    assert enc.precalculate()
    Z = enc.blocks()
    for rnd in range(3):
        for sbn in range(Z):
            assert enc.generate_symbols(sbn, io_in)
        enc.encoder_reset(0)
    packets = []
    for sbn in range(Z):  # dump_block :44-72
        Kb = enc.block_symbols(sbn)
        assert enc.generate_symbols(sbn, io_in)
        dropped = 0
        for esi in range(Kb):
            if rng.random() < loss:
                dropped += 1
            else:
                packets.append((api.tag(sbn, esi), enc.encode(esi, sbn, io_in).copy()))
        for esi in range(Kb, Kb + dropped + 2):
            packets.append((api.tag(sbn, esi), enc.encode(esi, sbn, io_in).copy()))
        enc.encoder_cleanup(sbn)
    out = np.zeros(F, np.uint8)
    io_out = nb.MemIO(out)
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    for rnd in range(3):
        out[:K * T] = 0  # block 0 is decoded afresh every round
        for t, r in packets:
            assert dec.add_symbol(r, int(t), io_out) != nb.SYM_ERR
        for sbn in range(Z):
            assert dec.repair_block(io_out, sbn), (rnd, sbn)
        assert np.array_equal(out, payload), rnd
        dec.encoder_reset(0)
    for sbn in range(Z):
        dec.encoder_cleanup(sbn)
    dec.close()
    enc.close()


def test_random_solver_requests_in_both_flavours():
    """rqb200.h level: random decode requests (repair ESIs anywhere, shuffled arrival, overhead 0..20) run with
    the flavour the planner picks and with the HBM flavour forced, including blocks batched into one launch;
    recovered symbols, intermediate symbols and the verdict against the oracle."""
    for seed in range(60):
        rng = np.random.default_rng(12000 + seed)
        K = int(rng.choice([rng.integers(1, 30), rng.integers(30, 300), rng.integers(300, 2500)]))
        T = int(rng.choice([8, 16, 40, 64, 1280]))
        p = orc_params(K)
        nblk = int(rng.integers(1, 4))
        jobs = []
        for b in range(nblk):
            src = rng.integers(0, 256, (K, T), dtype=np.uint8)
            Co, _, _ = orc_encode(K, T, src)
            drop = rng.random(K) < float(rng.choice([0.05, 0.3, 0.9]))
            drop[int(rng.integers(0, K))] = True
            rep = K + rng.choice(4 * K + 64, size=int(drop.sum()) + int(rng.choice([0, 1, 2, 20])), replace=False)
            esis = np.concatenate([np.nonzero(~drop)[0], rep]).astype(np.uint32)
            rng.shuffle(esis)
            syms = np.stack([src[e] if e < K else orc_lt(K, T, Co, int(e) + p.Kprime - K) for e in esis])
            rc_o, _, C_o, _, _ = orc_decode(K, T, esis, syms, want_C=True)
            req, missing = nb.SolveRequest.for_decoder(K, esis)
            jobs.append((src, esis, syms, rc_o, C_o, req, missing))
        n_in = max(len(j[1]) for j in jobs)
        for flavour in ("auto", "hbm"):
            solvers, live = [], []
            for src, esis, syms, rc_o, C_o, req, missing in jobs:
                s = nb.Solver(K, T, max_in=n_in, max_out=max(len(missing), 1), flavour=flavour)
                s.staging[:len(esis), :T] = syms
                s.upload(0, len(esis))
                rc = s.plan(req)
                assert (rc == 0) == (rc_o == 0), (seed, flavour, rc, rc_o)
                solvers.append(s)
                if rc == 0:
                    live.append(s)
            if len(live) > 1:
                nb.Solver.run_batch(live, live[0])
            elif live:
                live[0].run()
            for s, (src, esis, syms, rc_o, C_o, req, missing) in zip(solvers, jobs):
                if rc_o == 0:
                    assert np.array_equal(s.fetch_syms(len(missing)), src[missing]), (seed, flavour)
                    assert np.array_equal(s.fetch_c(), C_o), (seed, flavour)
                s.close()


def test_random_schedules_replayed_like_the_reference_applies_them():
    """rqb_schedule_replay / _stepwise on arbitrary op lists (not only the ones the reference's elimination
    emits): random axpy / scal sequences with long dependency chains and repeated destinations, random marks
    and row permutations; expected bytes = the literal order of precode_matrix_apply_sched + the swap walk of
    precode_matrix_permute (lib/precode.c:3-32,379-389), row ops by the oracle."""
    from oracle_lib import Op, oracle, ptr

    def oracle_ops(ops):
        o = (Op * len(ops))()
        for k, (b, i, j) in enumerate(zip(ops["beta"], ops["i"], ops["j"])):
            o[k].beta, o[k].i, o[k].j = int(b), int(i), int(j)
        return o

    def permute(D, P):
        P = list(P)
        for i in range(len(P)):
            at = i
            while P[at] >= 0:
                D[[i, P[at]]] = D[[P[at], i]]
                nxt = P[at]
                P[at] = -1
                at = nxt

    for seed in range(40):
        rng = np.random.default_rng(15000 + seed)
        R = int(rng.integers(2, 200))
        T = int(rng.choice([16, 48, 64, 1280]))
        nops = int(rng.integers(1, 6 * R))
        ops = np.zeros(nops, dtype=api.OP_DTYPE)
        hot = rng.integers(0, R, size=max(1, R // 8))  # a few rows take part in most ops: chains and repeats
        for k in range(nops):
            i = int(rng.choice(hot)) if rng.random() < 0.5 else int(rng.integers(0, R))
            if rng.random() < 0.85:
                j = int(rng.integers(0, R - 1))
                j += j >= i
                ops[k] = (int(rng.choice([1, 1, 1, rng.integers(2, 256)])), i, j)
            else:
                ops[k] = (0, i, int(rng.integers(1, 256)))
        m1 = int(rng.integers(1, nops + 1))
        m0 = int(rng.integers(0, m1))
        di = rng.permutation(R).astype(np.int32)
        c = rng.permutation(R).astype(np.int32)
        D = rng.integers(0, 256, (R, T), dtype=np.uint8)
        order = list(range(0, m1)) + list(range(m0, -1, -1)) + list(range(m1, nops)) + list(range(0, m0 + 1))
        seq = ops[order]
        want = D.copy()
        oracle().orc_apply_ops(ptr(want), T, T, oracle_ops(seq), len(seq))
        permute(want, di)
        permute(want, c)
        for stepwise in (False, True):
            m = nb.Matrix(R, T)
            m.upload(D)
            m.schedule_replay(ops, m0, m1, di, c, stepwise=stepwise)
            got = m.download()
            m.close()
            assert np.array_equal(got, want), (seed, stepwise, R, T, nops, m0, m1)


def test_calls_with_arguments_out_of_range_are_refused():
    """Block numbers beyond the object, ESIs beyond the range, short pitches, the wrong kind of object:
    every call says no (0 / false / NANORQ_SYM_ERR / -1) and the objects keep working afterwards."""
    K, T = 300, 64
    rng = np.random.default_rng(4)
    payload = rng.integers(0, 256, 2 * K * T, dtype=np.uint8)
    enc = nb.Encoder(len(payload), T, K, 0, 8)
    io_in = nb.MemIO(payload)
    Z = enc.blocks()
    assert Z == 2
    assert enc.generate_symbols(Z, io_in) is False
    assert enc.generate_symbols(255, io_in) is False
    assert enc.encode(0, Z, io_in) is None
    assert enc.encode(1 << 24, 0, io_in) is None
    assert enc.encode_range(Z, 0, 4, io_in) is None
    assert enc.encode_range(0, (1 << 24) - 2, 4, io_in) is None
    assert enc.block_symbols(Z) == 0
    wide = np.zeros((4, T - 8), np.uint8)
    assert nb.lib().nanorq_encode_range(enc.h, 0, 0, 4, wide.ctypes.data, T - 8, io_in.ptr) == 0  # pitch < T
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    out = np.zeros(len(payload), np.uint8)
    io_out = nb.MemIO(out)
    sym = enc.encode(0, 0, io_in)
    assert dec.add_symbol(sym, api.tag(Z, 0), io_out) == nb.SYM_ERR
    assert dec.add_symbol(sym, api.tag(0, 3 * K), io_out) == nb.SYM_ERR  # beyond max_esi
    assert dec.repair_block(io_out, Z) is False
    assert dec.num_missing(Z) == 0 and dec.num_repair(Z) == 0
    assert nb.lib().nanorq_encode_range(dec.h, 0, 0, 4, wide.ctypes.data, T, io_in.ptr) == 0  # a decoder object
    rows = np.zeros((4, T), np.uint8)
    tags = np.array([api.tag(0, 0)] * 4, np.uint32)
    assert nb.lib().nanorq_decoder_add_symbols(dec.h, tags.ctypes.data_as(api.u32p), rows.ctypes.data, T - 1, 4, None,
                                               io_out.ptr) == -1
    rc, st = dec.add_symbols(np.array([api.tag(Z, 0), api.tag(0, 0), api.tag(0, 3 * K)], np.uint32),
                             np.stack([sym, sym, sym]), io_out)
    assert rc == -1 and st == [nb.SYM_ERR, nb.SYM_ADDED, nb.SYM_ERR]
    assert dec.repair_blocks(io_out, [Z, 0, 200]) == [False, False, False]  # block 0 has one symbol so far
    # after all that: a normal transfer still works
    for sbn in range(Z):
        syms = enc.encode_range(sbn, 0, K + 7, io_in)
        tags = np.array([api.tag(sbn, e) for e in range(5, K + 7)], np.uint32)
        rc, _ = dec.add_symbols(tags, syms[5:], io_out)
        assert rc >= K - 6
        assert dec.repair_block(io_out, sbn)
    assert np.array_equal(out, payload)
    dec.close()
    enc.close()
