"""DMA from page-locked memory that is touched ONCE (freshly allocated arena, every block copied a
single time) against memory that is reused, and the same for DMA to host memory."""
import sys
import threading
import time
sys.path.insert(0, ".")
import nanorq_b200 as nb
K, T = 4096, 1280
L = nb.lib()


def run(nthr, blocks_per_thread, passes, mode):
    bufs = [nb.PinnedBuffer(blocks_per_thread * K * T) for _ in range(nthr)]
    for b in bufs:
        b.arr[:] = 1
    svs = [nb.Solver(K, T, max_in=K, max_out=K) for _ in range(nthr)]
    bar = threading.Barrier(nthr + 1)
    res = []
    def work(i):
        s, a = svs[i], bufs[i].arr
        for p_ in range(passes):
            bar.wait()
            for r in range(blocks_per_thread):
                p = a.ctypes.data + r * K * T
                if mode == "h2d":
                    L.rqb_solver_upload_rows(s.h, 0, K, p, T)
                else:
                    L.rqb_solver_fetch_rows(s.h, 0, 0, K, p, T, 0)
                L.rqb_solver_sync(s.h)
            bar.wait()
    th = [threading.Thread(target=work, args=(i,)) for i in range(nthr)]
    for t in th:
        t.start()
    for p_ in range(passes):
        bar.wait()
        t0 = time.perf_counter()
        bar.wait()
        res.append(nthr * blocks_per_thread * K * T / (time.perf_counter() - t0) / 1e9)
    for t in th:
        t.join()
    for s in svs:
        s.close()
    for b in bufs:
        b.close()
    return res


for mode in ("h2d", "d2h"):
    print(mode, "16 threads x 12 blocks of 5 MB, fresh arena, GB/s per pass:", ["%.1f" % v for v in run(16, 12, 3, mode)])
