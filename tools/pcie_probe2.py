"""Does DMA bandwidth depend on how much page-locked memory the copies walk over?  16 threads copy
5 MB blocks host->device and back, each thread cycling through its own arena of `span` MB."""
import sys
import threading
import time
sys.path.insert(0, ".")
import nanorq_b200 as nb
K, T = 4096, 1280
L = nb.lib()


def run(nthr, span_blocks, mode, reps=48):
    bufs = [nb.PinnedBuffer(span_blocks * K * T) for _ in range(nthr)]
    for b in bufs:
        b.arr[:] = 1
    svs = [nb.Solver(K, T, max_in=K, max_out=K) for _ in range(nthr)]
    bar = threading.Barrier(nthr + 1)

    def work(i):
        s, a = svs[i], bufs[i].arr
        bar.wait()
        for r in range(reps):
            p = a.ctypes.data + (r % span_blocks) * K * T
            if mode in ("h2d", "both"):
                L.rqb_solver_upload_rows(s.h, 0, K, p, T)
            if mode in ("d2h", "both"):
                L.rqb_solver_fetch_rows(s.h, 0, 0, K, p, T, 0)
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
    return nthr * reps * K * T / dt / 1e9


for span in (1, 4, 16):
    print("16 threads, %3d MB of pinned memory per thread: h2d %.1f, d2h %.1f, both %.1f GB/s each way" % (
        span * K * T >> 20, run(16, span, "h2d"), run(16, span, "d2h"), run(16, span, "both")))
