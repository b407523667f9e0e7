"""Concurrency stress of the nanorq.h layer on the GPU: several round-trip harness runs with
DIFFERENT block shapes at the same time (each with its own worker threads), so that solver
contexts of different sizes are recycled, encoder plans of different K are cached and built
concurrently, and many streams wait at once.  Every decoded byte is compared with the payload
inside the harness (bench/rq_roundtrip.c)."""
import ctypes as C
import os
import sys
import threading

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nanorq_b200 as nb  # noqa: E402

pytestmark = pytest.mark.gpu


def test_concurrent_roundtrips_of_different_shapes():
    L = C.CDLL(os.path.join(nb.api.LIB_DIR, "librq_roundtrip.so"))
    L.rq_roundtrip_run.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]
    # (K, T, loss, overhead, blocks, threads)
    shapes = [(10, 64, 0.3, 1, 300, 3), (257, 48, 0.2, 2, 120, 3), (1024, 1280, 0.05, 2, 48, 3),
              (1000, 520, 0.5, 1, 40, 2), (4096, 1280, 0.10, 0, 16, 3), (3000, 96, 0.9, 3, 12, 2)]
    results, errors = {}, []

    def run(idx, shape, rounds):
        K, T, loss, oh, nblocks, threads = shape
        for r in range(rounds):
            cfg = bench.RtConfig(K, T, nblocks, loss, oh, 17 * idx + r, threads, r % 2, 1)
            res = bench.RtResult()
            rc = L.rq_roundtrip_run(C.byref(cfg), C.byref(res))
            if rc != 0 or res.failures or res.mismatches:
                errors.append((shape, r, rc, res.failures, res.mismatches))
            results[(idx, r)] = res.out_fnv

    ts = [threading.Thread(target=run, args=(i, s, 3)) for i, s in enumerate(shapes)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    # the same seeds run alone give the same decoded bytes (digest over all blocks)
    for i, s in enumerate(shapes[:3]):
        K, T, loss, oh, nblocks, threads = s
        cfg = bench.RtConfig(K, T, nblocks, loss, oh, 17 * i, 1, 0, 1)
        res = bench.RtResult()
        assert L.rq_roundtrip_run(C.byref(cfg), C.byref(res)) == 0 and not res.failures and not res.mismatches
        assert res.out_fnv == results[(i, 0)]
    ev = nb.slow_path_counters()
    print("slow-path events over the stress run:", ev)
