#!/usr/bin/env python
"""End-to-end throughput of every BASELINE.json configuration: this library through nanorq.h
(bench/rq_roundtrip.c, per-symbol calls) and through nanorq_batch.h (bench/rq_roundtrip_batch.c,
page-locked buffers), next to the unmodified reference on the same source, seeds, thread count and
object shape (objects of ZBLOCKS blocks with nanorq_precalculate).  The decoded bytes of all three
are compared (FNV).  Not a bench line (bench.py reports C3 only): a table for DESIGN.md.

    python tools/config_table.py            (on a GPU box)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# name, K, T, loss, overhead, blocks per step, blocks per object
CONFIGS = [("C1", 10, 64, 0.0, 0, 16384, 16), ("C1-lossy", 10, 64, 0.3, 1, 8192, 16),
           ("C2", 1024, 1280, 0.05, 2, 512, 4), ("C3", 4096, 1280, 0.10, 0, 128, 4),
           ("C5", 56403, 512, 0.15, 0, 16, 1)]


def run(lib, fn, K, T, loss, oh, nblocks, threads, seed, z):
    L = C.CDLL(lib)
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]
    cfg = bench.RtConfig(K, T, nblocks, loss, oh, seed, threads, 1, 1, z)
    res = bench.RtResult()
    rc = f(C.byref(cfg), C.byref(res))
    assert rc == 0 and res.failures == 0 and res.mismatches == 0, (rc, res.failures, res.mismatches)
    return 2 * 8 * K * T * nblocks / res.wall_s / 1e9, res


def main():
    cores = os.cpu_count() or 1
    import nanorq_b200 as nbm
    own = os.path.join(ROOT, "nanorq_b200", "librq_roundtrip.so")
    batch = os.path.join(ROOT, "nanorq_b200", "librq_roundtrip_batch.so")
    ref = os.path.join(ROOT, "oracle", "_ref", "librq_roundtrip_ref.so")
    rows = []
    for name, K, T, loss, oh, nb, z in CONFIGS:
        th = max(1, min(nb // z, cores))
        for w in range(2):  # warm-up: plan caches, one context per thread and role
            run(own, "rq_roundtrip_run", K, T, loss, oh, nb, th, 5 + w, z)
            run(batch, "rq_roundtrip_batch_run", K, T, loss, oh, nb, th, 5 + w, z)
        g_own = max(run(own, "rq_roundtrip_run", K, T, loss, oh, nb, th, 1, z)[0] for _ in range(2))
        r_own = run(own, "rq_roundtrip_run", K, T, loss, oh, nb, th, 1, z)[1]
        g_bat = max(run(batch, "rq_roundtrip_batch_run", K, T, loss, oh, nb, th, 1, z)[0] for _ in range(2))
        r_bat = run(batch, "rq_roundtrip_batch_run", K, T, loss, oh, nb, th, 1, z)[1]
        g_ref, r_ref = run(ref, "rq_roundtrip_run", K, T, loss, oh, nb, th, 1, z)
        assert r_own.out_fnv == r_ref.out_fnv == r_bat.out_fnv, "decoded bytes differ between the builds"
        rows.append({"config": name, "K": K, "T": T, "loss": loss, "overhead": oh, "blocks": nb, "blocks_per_object": z,
                     "threads": th, "b200_per_symbol_gbit_s": round(g_own, 2), "b200_batch_gbit_s": round(g_bat, 2),
                     "reference_gbit_s": round(g_ref, 2), "ratio_per_symbol": round(g_own / g_ref, 2),
                     "ratio_batch": round(g_bat / g_ref, 2), "retries": r_own.retries})
        print(json.dumps(rows[-1]), flush=True)
    json.dump({"host_cores": cores, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "config_table.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
