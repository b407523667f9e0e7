"""Debug: per-level timeline of one CTA of the shared-memory solve kernel.
Needs a library built with RQB_NVCC_EXTRA=-DRQB_TRACE (tools/level_trace.sh)."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import nanorq_b200 as nb
from nanorq_b200 import workload
K, T = int(sys.argv[1]), int(sys.argv[2])
s = nb.Solver(K, T, max_in=K, max_out=16)
s.staging[:K, :T] = workload.payload(K, T, 1)
s.upload(0, K)
s.plan_encode(True, 0)
for r in range(3):
    s.run(); s.sync()
buf = (C.c_ulonglong * 16384)()
L = nb.lib()
L.rqb_dev_trace_fetch.argtypes = [C.POINTER(C.c_ulonglong), C.c_uint]
n = L.rqb_dev_trace_fetch(buf, 16384)
ev = [(int(v) >> 62, (int(v) >> 40) & 0x3FFFFF, int(v) & 0xFFFFFFFFFF) for v in buf[:n]]
t0 = ev[0][2]
print("events", n, "total cycles", ev[-1][2] - t0)
prev = t0
work = bar = page = 0
rows = []
for tag, val, clk in ev[1:]:
    d = clk - prev
    prev = clk
    if tag == 1: page += d; rows.append(("page", val, d))
    elif tag == 2: work += d; rows.append(("work", val, d))
    elif tag == 3: bar += d; rows.append(("bar", val, d))
print("thread0 cycles: in page waits %d, in level work %d, in barriers %d" % (page, work, bar))
# histogram of barrier-to-barrier time by task count
lv = []
cur = 0
for kind, val, d in rows:
    cur += d
    if kind == "bar":
        lv.append((val, cur)); cur = 0
lv = np.array(lv)
for lo, hi in ((0, 1), (1, 17), (17, 65), (65, 257), (257, 1025), (1025, 100000)):
    m = (lv[:, 0] >= lo) & (lv[:, 0] < hi)
    if m.any():
        print("levels with %5d..%5d tasks: %4d levels, mean %7.0f cycles, total %8d" % (lo, hi - 1, m.sum(), lv[m, 1].mean(), lv[m, 1].sum()))
print("first 60 level records (tasks, cycles):", [(int(a), int(b)) for a, b in lv[:60]])
pg = [d for k, v, d in rows if k == "page"]
print("page waits: n=%d mean %.0f max %d" % (len(pg), np.mean(pg), max(pg)))
