"""Where does a block's device pipeline saturate?  N threads, each with its own solver context, loop
over [upload K rows from page-locked memory, run the cached encode program, fetch n_out rows, wait].
Variants drop one stage at a time.  Prints blocks/s."""
import sys
import threading
import time
import numpy as np
sys.path.insert(0, ".")
import nanorq_b200 as nb
K, T = 4096, 1280
L = nb.lib()
NOUT = 512


def run(nthr, flavour, upload, kernel, fetch_rows, reps=30):
    bufs = [nb.PinnedBuffer((K + NOUT) * T) for _ in range(nthr)]
    svs = [nb.Solver(K, T, max_in=K, max_out=NOUT, flavour=flavour) for _ in range(nthr)]
    for s, b in zip(svs, bufs):
        L.rqb_solver_upload_rows(s.h, 0, K, b.arr.ctypes.data, T)
        s.plan_encode(True, NOUT)
        s.run(); s.sync()
    bar = threading.Barrier(nthr + 1)

    def work(i):
        s, a = svs[i], bufs[i].arr
        bar.wait()
        for r in range(reps):
            if upload:
                L.rqb_solver_upload_rows(s.h, 0, K, a.ctypes.data, T)
            if kernel:
                L.rqb_solver_run(s.h)
            if fetch_rows:
                L.rqb_solver_fetch_rows(s.h, 1, 0, fetch_rows, a.ctypes.data + K * T, T, 0)
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
    return nthr * reps / dt


for nthr in (8, 20):
    for fl in ("auto", "hbm"):
        print("threads %2d flavour %-4s: kernel only %6.0f | upload+kernel %6.0f | upload+kernel+fetch512 %6.0f | upload only %6.0f blocks/s" % (
            nthr, fl, run(nthr, fl, 0, 1, 0), run(nthr, fl, 1, 1, 0), run(nthr, fl, 1, 1, NOUT), run(nthr, fl, 1, 0, 0)))
