"""e2e round trips through nanorq.h (per-symbol arm, bench/rq_roundtrip.c) and through
nanorq_batch.h (batch arm, bench/rq_roundtrip_batch.c): K T loss overhead blocks threads [steps]."""
import ctypes as C
import os
import sys
sys.path.insert(0, ".")
import bench
import nanorq_b200 as nb
K, T, loss, oh, nblk, nthr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
steps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
nb.lib()
ARMS = (("per-symbol", "librq_roundtrip.so", "rq_roundtrip_run"), ("batch", "librq_roundtrip_batch.so", "rq_roundtrip_batch_run"))
if os.environ.get("XP_NOVERIFY"):
    ARMS = ARMS[1:]
for name, so, fn in ARMS:
    L = C.CDLL(os.path.join(nb.api.LIB_DIR, so))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(bench.RtConfig), C.POINTER(bench.RtResult)]
    vals = []
    for s in range(steps + 2):
        cfg = bench.RtConfig(K, T, nblk, loss, oh, 100 + s, nthr, 1, 0 if os.environ.get('XP_NOVERIFY') else 1, int(os.environ.get('ZBLOCKS', '1')))
        res = bench.RtResult()
        sp0 = nb.slow_path_counters()
        rc = f(C.byref(cfg), C.byref(res))
        sp1 = nb.slow_path_counters()
        slow = {k: sp1[k] - sp0[k] for k in sp1 if sp1[k] != sp0[k]}
        if s >= 2 and slow:
            print("   step %d slow-path events: %s" % (s, slow))
        assert rc == 0 and res.failures == 0 and res.mismatches == 0, (name, rc, res.failures, res.mismatches)
        if s >= 2:
            vals.append(2 * 8 * K * T * nblk / res.wall_s / 1e9)
    if os.environ.get("NANORQ_B200_PROFILE") == "1":
        prof = nb.host_profile(reset=True)
        tot = (steps + 2) * nblk
        print("   host profile, ms per block (thread time): " + ", ".join("%s %.2f" % (k, 1e3 * v / tot) for k, v in prof.items() if v > 0))
    print("%-10s K=%d T=%d: %s Gbit/s  (gen %.2f emit %.2f add %.2f repair %.2f ms/block thread time, fnv %x)" % (
        name, K, T, " ".join("%.1f" % v for v in vals), 1e3 * res.t_gen / nblk, 1e3 * res.t_emit / nblk,
        1e3 * res.t_add / nblk, 1e3 * res.t_repair / nblk, res.out_fnv))
