"""PCIe probe through the library's own copy paths: N threads, each with its own solver context,
copy K rows of T bytes host->device (rqb_solver_upload_rows from page-locked memory) and back
(rqb_solver_fetch_rows), waiting after each copy.  Prints aggregate GB/s per direction and both."""
import sys
import threading
import time
import numpy as np
sys.path.insert(0, ".")
import nanorq_b200 as nb
K, T = 4096, 1280
L = nb.lib()


def run(nthr, mode, reps=40):
    bufs = [nb.PinnedBuffer(K * T) for _ in range(nthr)]
    svs = [nb.Solver(K, T, max_in=K, max_out=K) for _ in range(nthr)]
    bar = threading.Barrier(nthr + 1)

    def work(i):
        s, a = svs[i], bufs[i].arr
        bar.wait()
        for r in range(reps):
            if mode in ("h2d", "both"):
                L.rqb_solver_upload_rows(s.h, 0, K, a.ctypes.data, T)
            if mode in ("d2h", "both"):
                L.rqb_solver_fetch_rows(s.h, 0, 0, K, a.ctypes.data, T, 0)
            L.rqb_solver_sync(s.h)
        bar.wait()

    th = [threading.Thread(target=work, args=(i,)) for i in range(nthr)]
    for t in th:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    for s in svs:
        s.close()
    for b in bufs:
        b.close()
    per_dir = nthr * reps * K * T / dt / 1e9
    return per_dir


for nthr in (1, 4, 16):
    print("threads %2d: h2d %.1f GB/s, d2h %.1f GB/s, both at once %.1f GB/s each way" % (
        nthr, run(nthr, "h2d"), run(nthr, "d2h"), run(nthr, "both")))
