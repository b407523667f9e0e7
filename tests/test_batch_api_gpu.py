"""nanorq_batch.h on the GPU: nanorq_encode_range / nanorq_decoder_add_symbols /
ioctx_from_pinned_mem / nanorq_set_devices against the per-symbol calls of nanorq.h and
against the oracle."""
import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import api
from oracle_lib import orc_encode, orc_lt, orc_params

pytestmark = pytest.mark.gpu


def pinned_array(nbytes):
    buf = nb.PinnedBuffer(nbytes)
    return buf, buf.arr


@pytest.mark.parametrize("F,T,K", [(640, 64, 10), (100 * 64 - 17, 64, 100), (1310720, 1280, 1024), (5242880, 1280, 4096)])
@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_encode_range_equals_per_symbol_encode_and_oracle(F, T, K, pinned):
    rng = np.random.default_rng(F)
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    keep = []
    if pinned:
        buf, arr = pinned_array(F)
        arr[:] = payload
        keep.append(buf)
        io = nb.PinnedMemIO(arr)
    else:
        io = nb.MemIO(payload)
    enc = nb.Encoder(F, T, K, 0, 8)
    ref = nb.Encoder(F, T, K, 0, 8)
    io_ref = nb.MemIO(payload)
    Kb = enc.block_symbols(0)
    # a range that straddles source and repair symbols and more than one repair window
    n_rep = max(40, Kb // 8 + 70)
    first = max(0, Kb - 7)
    if pinned:
        obuf, oarr = pinned_array((Kb + n_rep) * (T + 16))
        keep.append(obuf)
        out = oarr.reshape(Kb + n_rep, T + 16)  # rows wider than T: the pitch is honoured
    else:
        out = np.zeros((Kb + n_rep, T), np.uint8)
    got = enc.encode_range(0, first, Kb - first + n_rep, io, out=out)
    assert got is not None
    for k, esi in enumerate(range(first, Kb + n_rep)):
        assert np.array_equal(got[k], ref.encode(esi, 0, io_ref)), esi
    # all source symbols in one call, then an isolated far repair range
    src = enc.encode_range(0, 0, Kb, io)
    blk = np.zeros(Kb * T, np.uint8)
    blk[:F] = payload
    assert np.array_equal(src, blk.reshape(Kb, T))
    far = enc.encode_range(0, Kb + 5000, 9, io)
    p = orc_params(Kb)
    Co, _, _ = orc_encode(Kb, T, blk)
    for k in range(9):
        assert np.array_equal(far[k], orc_lt(Kb, T, Co, Kb + 5000 + k + p.Kprime - Kb))
    # per-symbol calls still work after range calls (window bookkeeping)
    assert np.array_equal(enc.encode(Kb + 3, 0, io), orc_lt(Kb, T, Co, Kb + 3 + p.Kprime - Kb))
    enc.close(); ref.close(); io.close(); io_ref.close()
    for b in keep:
        b.close()


def make_packets(F, T, K, Z, loss, oh, seed):
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    enc = nb.Encoder(F, T, K, Z, 8)
    io = nb.MemIO(payload)
    tags, rows = [], []
    for sbn in range(enc.blocks()):
        Kb = enc.block_symbols(sbn)
        drop = rng.random(Kb) < loss
        n_rep = int(drop.sum()) + oh
        syms = enc.encode_range(sbn, 0, Kb + n_rep, io)
        for esi in list(np.nonzero(~drop)[0]) + list(range(Kb, Kb + n_rep)):
            tags.append(api.tag(sbn, int(esi)))
            rows.append(syms[esi])
    oti = (enc.oti_common(), enc.oti_scheme_specific())
    enc.close()
    return payload, oti, np.array(tags, np.uint32), np.stack(rows)


@pytest.mark.parametrize("F,T,K,Z,loss,oh", [(640, 64, 10, 0, 0.3, 2), (100 * 64 - 17, 64, 100, 0, 0.2, 2),
                                              (1310720, 1280, 1024, 0, 0.05, 2), (5242880, 1280, 4096, 0, 0.10, 2),
                                              (3 * 500 * 104 - 5, 104, 500, 0, 0.15, 3)])
@pytest.mark.parametrize("mode", ["pageable", "pinned", "pinned-shuffled", "mixed-calls"])
def test_add_symbols_decodes_like_per_symbol_calls(F, T, K, Z, loss, oh, mode):
    payload, oti, tags, rows = make_packets(F, T, K, Z, loss, oh, seed=F % 89 + 1)
    keep = []
    if mode == "pinned-shuffled":
        order = np.random.default_rng(1).permutation(len(tags))
        tags, rows = tags[order], rows[order]
    if mode.startswith("pinned"):
        pb, parr = pinned_array(rows.size)
        parr[:] = rows.reshape(-1)
        data = parr.reshape(rows.shape)
        ob, out = pinned_array(F)
        out[:] = 0xEE
        keep += [pb, ob]
        io = nb.PinnedMemIO(out)
    else:
        data = rows
        out = np.full(F, 0xEE, np.uint8)
        io = nb.MemIO(out)
    dec = nb.Decoder(*oti)
    if mode == "mixed-calls":  # alternate between the batch call and the per-symbol call on the same blocks
        k = 0
        while k < len(tags):
            n = min(37, len(tags) - k)
            rc, st = dec.add_symbols(tags[k:k + n], data[k:k + n], io)
            assert rc >= 0
            k += n
            for q in range(k, min(k + 5, len(tags))):
                assert dec.add_symbol(data[q], int(tags[q]), io) in (nb.SYM_ADDED, nb.SYM_IGN)
            k = min(k + 5, len(tags))
    else:
        rc, st = dec.add_symbols(tags, data, io)
        assert rc >= 0 and all(s in (nb.SYM_ADDED, nb.SYM_IGN) for s in st)
        # a second delivery of the same symbols is classified like the per-symbol call does
        rc2, st2 = dec.add_symbols(tags[:5], data[:5], io)
        assert rc2 == 0 and all(s in (nb.SYM_DUP, nb.SYM_IGN) for s in st2)
    for sbn in range(dec.blocks()):
        assert dec.repair_block(io, sbn), sbn
        assert dec.repair_block(io, sbn)  # idempotent
    assert np.array_equal(out, payload)
    dec.close(); io.close()
    for b in keep:
        b.close()


def test_pinned_output_block_completed_by_the_last_source_symbol():
    """No loss: the call that delivers a block's last source symbol hands the block back."""
    F, T, K = 300 * 48, 48, 300
    payload, oti, tags, rows = make_packets(F, T, K, 0, 0.0, 0, seed=5)
    pb, parr = pinned_array(rows.size)
    parr[:] = rows.reshape(-1)
    ob, out = pinned_array(F)
    out[:] = 0
    io = nb.PinnedMemIO(out)
    dec = nb.Decoder(*oti)
    rc, _ = dec.add_symbols(tags[:200], parr.reshape(rows.shape)[:200], io)
    assert rc == 200 and dec.num_missing(0) == 100
    rc, _ = dec.add_symbols(tags[200:], parr.reshape(rows.shape)[200:], io)
    assert rc == 100 and dec.num_missing(0) == 0
    assert np.array_equal(out, payload)
    assert dec.repair_block(io, 0)
    dec.close(); io.close(); pb.close(); ob.close()


def test_decoder_reset_with_deferred_output_decodes_again():
    """nanorq_encoder_reset (nanorq.h:79) on a decoder block whose output is handed back as a block
    image: the second round is written too, into whatever ioctx it comes with."""
    F, T, K = 600 * 1280, 1280, 600
    dec = None
    for rnd, (loss, seed) in enumerate([(0.0, 21), (0.1, 22)]):
        payload, oti, tags, rows = make_packets(F, T, K, 0, loss, 2, seed=seed)
        pb, parr = pinned_array(rows.size)
        parr[:] = rows.reshape(-1)
        ob, out = pinned_array(F)
        out[:] = 0
        io = nb.PinnedMemIO(out)
        if dec is None:
            dec = nb.Decoder(*oti)
        else:
            dec.encoder_reset(0)
            assert dec.num_missing(0) == K and dec.num_repair(0) == 0
        rc, _ = dec.add_symbols(tags, parr.reshape(rows.shape), io)
        assert rc == (K if loss == 0.0 else len(tags))  # symbols after completion are ignored
        assert dec.repair_block(io, 0)
        assert np.array_equal(out, payload), rnd
        io.close(); pb.close(); ob.close()
    dec.close()


def test_pinned_ioctx_with_the_per_symbol_api_only():
    """ioctx_from_pinned_mem behind the unchanged nanorq.h calls: load by DMA, output deferred."""
    F, T, K = 1000 * 200 - 33, 200, 1000
    rng = np.random.default_rng(8)
    ib, inp = pinned_array(F)
    inp[:] = rng.integers(0, 256, F, dtype=np.uint8)
    ob, out = pinned_array(F)
    out[:] = 0
    io_in, io_out = nb.PinnedMemIO(inp), nb.PinnedMemIO(out)
    enc = nb.Encoder(F, T, K, 0, 8)
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    assert enc.generate_symbols(0, io_in)
    drop = rng.random(K) < 0.1
    for esi in list(np.nonzero(~drop)[0]) + list(range(K, K + int(drop.sum()) + 2)):
        assert dec.add_symbol(enc.encode(int(esi), 0, io_in), api.tag(0, int(esi)), io_out) == nb.SYM_ADDED
    assert dec.repair_block(io_out, 0)
    assert np.array_equal(out, inp)
    # a memory the library page-locks itself
    plain = np.array(inp)
    io2 = nb.PinnedMemIO(plain, already_pinned=False)
    enc2 = nb.Encoder(F, T, K, 0, 8)
    assert np.array_equal(enc2.encode_range(0, K, 4, io2), enc.encode_range(0, K, 4, io_in))
    for x in (enc, enc2, dec, io_in, io_out, io2, ib, ob):
        x.close()


def test_blocks_of_one_object_spread_over_all_devices():
    """nanorq_set_devices(0): block sbn is solved on device sbn mod n; one process, one object."""
    n = nb.device_count()
    F, T, K, Z = 8 * 1024 * 256, 256, 1024, 8
    payload, oti, tags, rows = make_packets(F, T, K, 0, 0.1, 2, seed=12)
    dec = nb.Decoder(*oti)
    assert dec.set_devices(0) == n
    out = np.zeros(F, np.uint8)
    io = nb.MemIO(out)
    rc, _ = dec.add_symbols(tags, rows, io)
    assert rc >= 0
    for sbn in range(dec.blocks()):
        assert dec.repair_block(io, sbn)
    assert np.array_equal(out, payload)
    enc = nb.Encoder(F, T, K, 0, 8)
    assert enc.set_devices(0) == n
    io_in = nb.MemIO(payload)
    p = orc_params(K)
    for sbn in (0, 1, Z - 1):
        Co, _, _ = orc_encode(K, T, payload[sbn * K * T:(sbn + 1) * K * T])
        got = enc.encode_range(sbn, K, 3, io_in)
        assert np.array_equal(got, np.stack([orc_lt(K, T, Co, K + k + p.Kprime - K) for k in range(3)]))
    print("devices used:", n)
    for x in (enc, dec, io, io_in):
        x.close()


def test_solver_on_every_device_and_flavour_switch():
    K, T = 500, 64
    rng = np.random.default_rng(4)
    src = rng.integers(0, 256, (K, T), dtype=np.uint8)
    Co, _, _ = orc_encode(K, T, src)
    for dev in range(nb.device_count()):
        for fl in ("auto", "hbm"):
            s = nb.Solver(K, T, max_in=K, max_out=4, device=dev, flavour=fl)
            s.staging[:K, :T] = src
            s.upload(0, K)
            s.plan_encode(True, 0)
            assert bool(s.stats()["smem"]) == (fl == "auto")
            s.run()
            assert np.array_equal(s.fetch_c(), Co), (dev, fl)
            s.close()


def test_receive_ring_with_holes_is_added_with_one_call():
    """NANORQ_TAG_NONE marks rows of the caller's buffer that hold no symbol (lost packets in a
    ring indexed by sequence number); everything else in the ring is added by one call."""
    F, T, K = 700 * 48 - 5, 48, 700
    rng = np.random.default_rng(21)
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    enc = nb.Encoder(F, T, K, 0, 8)
    io_in = nb.MemIO(payload)
    drop = rng.random(K) < 0.25
    n_rep = int(drop.sum()) + 3
    pb, parr = pinned_array((K + n_rep) * T)
    ring = parr.reshape(K + n_rep, T)
    assert enc.encode_range(0, 0, K + n_rep, io_in, out=ring) is not None
    tags = np.array([0xFFFFFFFF if (e < K and drop[e]) else api.tag(0, e) for e in range(K + n_rep)], np.uint32)
    ring[np.nonzero(drop)[0]] = 0x77  # whatever is in a hole must not matter
    ob, out = pinned_array(F)
    io_out = nb.PinnedMemIO(out)
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    rc, st = dec.add_symbols(tags, ring, io_out)
    assert rc == K - int(drop.sum()) + n_rep
    assert all(s == nb.SYM_IGN for s, t in zip(st, tags) if t == 0xFFFFFFFF)
    assert dec.repair_block(io_out, 0)
    assert np.array_equal(out, payload)
    for x in (enc, dec, io_in, io_out, pb, ob):
        x.close()


@pytest.mark.parametrize("pinned_out", [False, True], ids=["pageable-out", "pinned-out"])
def test_repair_blocks_one_launch_for_all_blocks_of_an_object(pinned_out):
    """nanorq_repair_blocks: the 8 blocks of a C4-shaped object are repaired with one solve launch
    (per device); one block is short of symbols and must come back false, the others true."""
    F, T, K = 8 * 1024 * 320, 320, 1024
    payload, oti, tags, rows = make_packets(F, T, K, 0, 0.1, 2, seed=31)
    # starve block 5: drop its repair symbols
    keep = np.array([not ((int(t) >> 24) == 5 and (int(t) & 0xFFFFFF) >= K) for t in tags])
    tags2, rows2 = tags[keep], rows[keep]
    keepalive = []
    if pinned_out:
        ob, out = pinned_array(F)
        out[:] = 0
        keepalive.append(ob)
        io = nb.PinnedMemIO(out)
    else:
        out = np.zeros(F, np.uint8)
        io = nb.MemIO(out)
    dec = nb.Decoder(*oti)
    rc, _ = dec.add_symbols(tags2, rows2, io)
    assert rc >= 0
    l0 = nb.kernel_launches()
    res = dec.repair_blocks(io, list(range(8)))
    solves = nb.kernel_launches() - l0
    assert res == [True] * 5 + [False] + [True] * 2
    # ONE solve launch for the seven decodable blocks (+ one row-placement kernel per block with deferred output)
    assert solves == (8 if pinned_out else 1), solves
    # the starved block decodes once its repair symbols arrive; repeated calls are idempotent
    late = np.array([(int(t) >> 24) == 5 and (int(t) & 0xFFFFFF) >= K for t in tags])
    rc, _ = dec.add_symbols(tags[late], rows[late], io)
    assert rc > 0
    assert dec.repair_blocks(io, [5, 0, 3]) == [True, True, True]
    assert np.array_equal(out, payload)
    # a block listed twice is repaired once
    dec2 = nb.Decoder(*oti)
    out2 = np.zeros(F, np.uint8)
    io2 = nb.MemIO(out2)
    assert dec2.add_symbols(tags, rows, io2)[0] >= 0
    assert dec2.repair_blocks(io2, [1, 1, 2, 1, 0, 3, 4, 5, 6, 7, 7]) == [True] * 11
    assert np.array_equal(out2, payload)
    dec2.close(); io2.close()
    dec.close(); io.close()
    for b in keepalive:
        b.close()


def test_plan_batch_on_several_threads_gives_the_same_results():
    """rqb_solver_plan_batch: 24 decode blocks planned on 6 host threads, one batched launch, results
    against the source; a singular request in the batch comes back as 1 without disturbing the others."""
    from nanorq_b200 import workload
    K, T, n = 1024, 256, 24
    p = nb.block_params(K)
    srcs, decs, reqs, miss = [], [], [], []
    for b in range(n):
        src = workload.payload(K, T, 50 + b)
        e = nb.Solver(K, T, max_in=K, max_out=128)
        e.staging[:K, :T] = src
        e.upload(0, K)
        e.plan_encode(True, 128)
        e.run()
        rep = e.fetch_syms(128)
        e.close()
        drop = workload.loss_pattern(K, 0.08, b)
        esis = workload.received_esis(K, drop, 2, 0)
        req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
        if b == 7:  # make this one singular: one repair symbol stands in for four missing source symbols
            isi = req.isi.copy()           # (two more equal rows than the overhead of 2 can make up for)
            isi[missing[1]] = isi[missing[2]] = isi[missing[3]] = isi[missing[0]]
            req = nb.SolveRequest(isi, req.in_row, req.c.overhead, False, missing)
        d = nb.Solver(K, T, max_in=len(esis), max_out=len(missing))
        d.staging[:len(esis), :T] = np.concatenate([src[~drop], rep[:len(esis) - int((~drop).sum())]])
        d.upload(0, len(esis))
        srcs.append(src); decs.append(d); reqs.append(req); miss.append(missing)
    rcs = nb.Solver.plan_batch(decs, reqs, 6)
    assert rcs == [0] * 7 + [1] + [0] * 16
    good = [d for k, d in enumerate(decs) if rcs[k] == 0]
    nb.Solver.run_batch(good, good[0])
    for k, d in enumerate(decs):
        if rcs[k] == 0:
            assert np.array_equal(d.fetch_syms(len(miss[k])), srcs[k][miss[k]]), k
        d.close()
    # and through nanorq_repair_blocks with planning threads switched on
    nb.lib().rqb_set_plan_threads(4)
    try:
        F, T2, K2 = 8 * 512 * 128, 128, 512
        payload, oti, tags, rows = make_packets(F, T2, K2, 0, 0.1, 2, seed=77)
        dec = nb.Decoder(*oti)
        out = np.zeros(F, np.uint8)
        io = nb.MemIO(out)
        assert dec.add_symbols(tags, rows, io)[0] >= 0
        assert dec.repair_blocks(io, list(range(8))) == [True] * 8
        assert np.array_equal(out, payload)
        dec.close(); io.close()
    finally:
        nb.lib().rqb_set_plan_threads(1)
