#!/usr/bin/env python
"""Host time accounting of the nanorq.h layer for one configuration:
   NANORQ_B200_PROFILE=1 python tools/profile_config.py K T loss overhead blocks [threads]"""
import ctypes as C
import os
import sys
os.environ["NANORQ_B200_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import nanorq_b200 as nbm

K, T, loss, oh, nb = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
threads = int(sys.argv[6]) if len(sys.argv) > 6 else (5 * (os.cpu_count() or 1)) // 4
L = C.CDLL(os.path.join(nbm.api.LIB_DIR, "librq_roundtrip.so"))
L.rq_roundtrip_run.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]


def run(seed):
    cfg = bench.RtConfig(K, T, nb, loss, oh, seed, threads, 1, 1)
    res = bench.RtResult()
    assert L.rq_roundtrip_run(C.byref(cfg), C.byref(res)) == 0 and not res.failures and not res.mismatches
    return res


run(1); run(2)
nbm.host_profile(reset=True)
l0 = nbm.kernel_launches()
r = run(3)
print("K=%d T=%d blocks=%d threads=%d: %.2f Gbit/s, %.1f us wall per block, %.1f launches per block" % (
    K, T, nb, threads, 2 * 8 * K * T * nb / r.wall_s / 1e9, 1e6 * r.wall_s / nb, (nbm.kernel_launches() - l0) / nb))
print("  harness phases us/block:", [round(1e6 * x / nb, 1) for x in (r.t_gen, r.t_emit, r.t_add, r.t_repair)])
print("  library us/block:", {k: round(1e6 * v / nb, 1) for k, v in nbm.host_profile().items() if v > 0})
