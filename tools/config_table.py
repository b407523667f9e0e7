#!/usr/bin/env python
"""End-to-end throughput of every BASELINE.json configuration through nanorq.h
(bench/rq_roundtrip.c), this library next to the unmodified reference, same seeds.
Not a bench line (bench.py reports C3 only): a table for DESIGN.md.

    python tools/config_table.py            (on a GPU box)"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CONFIGS = [("C1", 10, 64, 0.0, 0, 4096), ("C2", 1024, 1280, 0.05, 2, 256), ("C3", 4096, 1280, 0.10, 0, 96),
           ("C5", 56403, 512, 0.15, 0, 16)]


def run(lib, K, T, loss, oh, nblocks, threads, seed, precalc):
    L = C.CDLL(lib)
    L.rq_roundtrip_run.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]
    cfg = bench.RtConfig(K, T, nblocks, loss, oh, seed, threads, precalc, 1)
    res = bench.RtResult()
    rc = L.rq_roundtrip_run(C.byref(cfg), C.byref(res))
    assert rc == 0 and res.failures == 0 and res.mismatches == 0, (rc, res.failures, res.mismatches)
    return 2 * 8 * K * T * nblocks / res.wall_s / 1e9, res


def main():
    cores = os.cpu_count() or 1
    own = os.path.join(ROOT, "nanorq_b200", "librq_roundtrip.so")
    ref = os.path.join(ROOT, "oracle", "_ref", "librq_roundtrip_ref.so")
    rows = []
    for name, K, T, loss, oh, nb in CONFIGS:
        th_own = max(1, min(nb, (5 * cores) // 4))
        for w in range(2):  # warm-up: plan caches, one context per thread and role
            run(own, K, T, loss, oh, min(nb, 3 * th_own), th_own, 5 + w, 1)
        import nanorq_b200 as nbm
        nbm.host_profile(reset=True)
        g_own, r_own = run(own, K, T, loss, oh, nb, th_own, 1, 1)
        if os.environ.get("NANORQ_B200_PROFILE") == "1":
            print("   host ms/block:", {k: round(1e3 * v / nb, 3) for k, v in nbm.host_profile().items() if v > 0},
                  "phases", [round(1e3 * x / nb, 2) for x in (r_own.t_gen, r_own.t_emit, r_own.t_add, r_own.t_repair)])
        g_ref, r_ref = run(ref, K, T, loss, oh, nb, min(nb, cores), 1, 0)
        assert r_own.out_fnv == r_ref.out_fnv, "decoded bytes differ between the two builds"
        rows.append({"config": name, "K": K, "T": T, "loss": loss, "overhead": oh, "blocks": nb,
                     "b200_gbit_s": round(g_own, 2), "reference_gbit_s": round(g_ref, 2), "ratio": round(g_own / g_ref, 2),
                     "threads_b200": th_own, "threads_reference": min(nb, cores), "retries": r_own.retries})
        print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_table.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
