#!/usr/bin/env python
"""e2e arm only, for one build of the library:  python tools/ab_e2e.py <libdir> [threads] [steps] [K T loss overhead blocks]
(A/B comparisons of two builds in one GPU session; loads nothing but the round-trip harness)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

libdir = sys.argv[1]
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
K, T, loss, oh, nb = (int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8])) if len(sys.argv) > 8 else (4096, 1280, 0.10, 0, 118)
L = C.CDLL(os.path.join(libdir, "librq_roundtrip.so"))
L.rq_roundtrip_run.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]


def run(nblocks, seed):
    cfg = bench.RtConfig(K, T, nblocks, loss, oh, seed, threads, 1, 1)
    res = bench.RtResult()
    rc = L.rq_roundtrip_run(C.byref(cfg), C.byref(res))
    assert rc == 0 and not res.failures and not res.mismatches
    return res


N = C.CDLL(os.path.join(libdir, "libnanorq_b200.so"))


def counters():
    out = (C.c_ulonglong * 4)()
    try:
        N.rqb_slow_path_counters(C.byref(out))
    except AttributeError:
        return None
    return list(out)


for w in range(3):
    run(min(nb, 2 * threads), 900 + w)
c0 = counters()
import resource
ru0 = resource.getrusage(resource.RUSAGE_SELF)
tot, parts = 0.0, [0.0] * 4
per = []
for s in range(steps):
    r = run(nb, s)
    tot += r.wall_s
    per.append(round(2 * 8 * K * T * nb / r.wall_s / 1e9, 1))
    for k, v in enumerate((r.t_gen, r.t_emit, r.t_add, r.t_repair)):
        parts[k] += v
ru1 = resource.getrusage(resource.RUSAGE_SELF)
print("   CPU per block: user %.2f ms, sys %.2f ms; wall x cores per block %.2f ms (%d cores)" % (
    1e3 * (ru1.ru_utime - ru0.ru_utime) / (nb * steps), 1e3 * (ru1.ru_stime - ru0.ru_stime) / (nb * steps),
    1e3 * tot * (os.cpu_count() or 1) / (nb * steps), os.cpu_count() or 1),
    "| minor faults per block %.0f, vol/invol ctx switches per block %.1f/%.1f" % (
    (ru1.ru_minflt - ru0.ru_minflt) / (nb * steps), (ru1.ru_nvcsw - ru0.ru_nvcsw) / (nb * steps), (ru1.ru_nivcsw - ru0.ru_nivcsw) / (nb * steps)))
print("   slow-path events during the timed steps {pinned, device, regrow, contexts}:", None if c0 is None else [a - b for a, b in zip(counters(), c0)])
print("%s: %.1f Gbit/s (per step %s) phases ms/block %s" % (os.path.basename(libdir.rstrip("/")), 2 * 8 * K * T * nb * steps / tot / 1e9, per,
                                                       [round(1e3 * x / (nb * steps), 2) for x in parts]))
