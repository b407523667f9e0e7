"""Single-block solve latency (device time, CUDA events): K T [repeats]."""
import sys
import numpy as np
sys.path.insert(0, ".")
import nanorq_b200 as nb
from nanorq_b200 import workload
K, T = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
s = nb.Solver(K, T, max_in=K, max_out=16)
s.staging[:K, :T] = workload.payload(K, T, 1)
s.upload(0, K)
s.plan_encode(True, 0)
ms = []
for r in range(reps):
    s.run(); s.sync(); ms.append(s.last_kernel_ms())
st = s.stats()
print("K=%d T=%d encode: first %.3f ms, median of rest %.3f ms, levels %d tasks %d -> %.0f ns/level" % (
    K, T, ms[0], float(np.median(ms[1:])), st["n_levels"], st["n_tasks"], 1e6 * float(np.median(ms[1:])) / st["n_levels"]))
drop = workload.loss_pattern(K, 0.1, 3)
esis = workload.received_esis(K, drop, 0)
req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)  # as nanorq_repair_block does
d = nb.Solver(K, T, max_in=len(esis), max_out=len(missing))
d.staging[:len(esis), :T] = 7
d.upload(0, len(esis))
assert d.plan(req) == 0
ms = []
for r in range(reps):
    d.run(); d.sync(); ms.append(d.last_kernel_ms())
st = d.stats()
print("K=%d T=%d decode: first %.3f ms, median of rest %.3f ms, levels %d tasks %d -> %.0f ns/level" % (
    K, T, ms[0], float(np.median(ms[1:])), st["n_levels"], st["n_tasks"], 1e6 * float(np.median(ms[1:])) / st["n_levels"]))
