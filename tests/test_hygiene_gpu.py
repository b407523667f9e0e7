"""Resource hygiene of the device context cache (rqb_solver.c) and ordering of batched
launches across streams, on the GPU."""
import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import api
from oracle_lib import orc_encode, orc_lt, orc_params

pytestmark = pytest.mark.gpu


def roundtrip_object(F, T, K, Z, loss, seed):
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 256, F, dtype=np.uint8)
    enc = nb.Encoder(F, T, K, Z, 8)
    io_in = nb.MemIO(payload)
    dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
    out = np.zeros(F, np.uint8)
    io_out = nb.MemIO(out)
    for sbn in range(enc.blocks()):
        Kb = enc.block_symbols(sbn)
        assert enc.generate_symbols(sbn, io_in)
        drop = rng.random(Kb) < loss
        esis = list(np.nonzero(~drop)[0]) + list(range(Kb, Kb + int(drop.sum()) + 2))
        for e in esis:
            dec.add_symbol(enc.encode(int(e), sbn, io_in), api.tag(sbn, int(e)), io_out)
        assert dec.repair_block(io_out, sbn)
    enc.close()
    dec.close()
    assert np.array_equal(out, payload)


def test_soak_1000_objects_of_mixed_shapes_keeps_memory_flat():
    """1000 objects of 40 different (K, T) shapes come and go under a small cache limit:
    the bytes parked by the library stay under the limit and the device's free memory
    does not drift (everything beyond the limit really goes back to the driver)."""
    nb.release_cached()
    nb.set_cache_limit(96 << 20)
    try:
        rng = np.random.default_rng(9)
        shapes = [(int(rng.integers(3, 600)), int(rng.choice([8, 24, 64, 104, 256, 1000, 1280]))) for _ in range(40)]
        free_marks = []
        for k in range(1000):
            K, T = shapes[int(rng.integers(0, len(shapes)))]
            roundtrip_object(K * T - int(rng.integers(0, T)), T, K, 0, 0.2, k)
            cached, limit = nb.cache_stats()
            assert cached <= limit + (64 << 20), (k, cached, limit)  # one context may be parked before the trim
            if k % 100 == 99:
                free_marks.append(nb.device_mem_info()[0])
        print("device free MB at every 100th object:", [f >> 20 for f in free_marks])
        # after warm-up the device's free memory must not shrink by more than the cache limit
        assert free_marks[1] - free_marks[-1] <= (160 << 20), free_marks
        nb.release_cached()
        assert nb.cache_stats()[0] == 0
    finally:
        nb.set_cache_limit(48 << 30)


def test_release_cached_returns_memory_to_the_driver():
    nb.release_cached()
    free0 = nb.device_mem_info()[0]
    for seed in range(3):
        roundtrip_object(1024 * 1280, 1280, 1024, 0, 0.1, seed)
    assert nb.cache_stats()[0] > 0
    nb.release_cached()
    assert nb.cache_stats()[0] == 0
    free1 = nb.device_mem_info()[0]
    assert free0 - free1 <= (8 << 20), (free0, free1)  # nothing but allocator granularity is left behind


def test_batched_launch_is_ordered_with_the_members_own_streams():
    """rqb_solver_run_batch launches on the owner's stream; a member's fetch, emit or next upload
    is queued on the member's own stream right away, with no host synchronisation in between,
    and must still see the batched kernel's results (ordering by events on the device)."""
    K, T, n = 1024, 1280, 12
    p = orc_params(K)
    rng = np.random.default_rng(3)
    srcs = [rng.integers(0, 256, (K, T), dtype=np.uint8) for _ in range(n)]
    for rep in range(3):
        encs = []
        for b in range(n):
            e = nb.Solver(K, T, max_in=K, max_out=32)
            e.staging[:K, :T] = srcs[b]
            e.upload(0, K)  # still queued on the member's stream when the batch is launched
            e.plan_encode(True, 16)
            encs.append(e)
        nb.Solver.run_batch(encs, encs[0])
        # no sync of the owner: members go on immediately on their own streams
        isi = np.arange(p.Kprime + 16, p.Kprime + 32, dtype=np.uint32)
        outs = []
        for b, e in enumerate(encs[::-1]):
            first = e.fetch_syms(16)          # symbols emitted with the solve
            e.emit(isi)                       # LT kernel reading C on the member's stream
            outs.append((first, e.fetch_syms(16)))
        for b, e in enumerate(encs[::-1]):
            Co, _, _ = orc_encode(K, T, srcs[n - 1 - b])
            first, second = outs[b]
            assert np.array_equal(first, np.stack([orc_lt(K, T, Co, p.Kprime + k) for k in range(16)]))
            assert np.array_equal(second, np.stack([orc_lt(K, T, Co, int(x)) for x in isi]))
            e.close()


def test_one_owner_launches_two_different_batches_back_to_back():
    """The owner lends its argument buffer to a batched launch; a second batch with OTHER blocks on
    the same owner, queued while the first may not even have started, must not disturb the first."""
    K, T, n = 1024, 1280, 8
    p = orc_params(K)
    rng = np.random.default_rng(8)
    srcs = [rng.integers(0, 256, (K, T), dtype=np.uint8) for _ in range(2 * n)]
    encs = []
    for b in range(2 * n):
        e = nb.Solver(K, T, max_in=K, max_out=32)
        e.staging[:K, :T] = srcs[b]
        e.upload(0, K)
        e.plan_encode(True, 16)
        encs.append(e)
    own = encs[0]
    for rep in range(4):  # something long on the owner's stream first, so that the copies queue up behind it
        nb.Solver.run_batch(encs[:n], own)
    nb.Solver.run_batch([own] + encs[n:], own)
    nb.Solver.run_batch(encs[:n], own)
    for b, e in enumerate(encs):
        got = e.fetch_syms(16)
        Co, _, _ = orc_encode(K, T, srcs[b])
        assert np.array_equal(got, np.stack([orc_lt(K, T, Co, p.Kprime + k) for k in range(16)])), b
    for e in encs:
        e.close()
