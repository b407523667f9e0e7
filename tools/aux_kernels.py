"""Exercises the two kernels bench.py's kernel arm does not launch, for the ncu capture
in tools/profile_solve.sh: rqb_lt_kernel (on-demand LT combine of repair symbols) and
rqb_gather_rows_kernel + per-level rqb_rowops_kernel (reference-schedule replay)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import nanorq_b200 as nb
from nanorq_b200 import workload

K, T, N = 4096, 1280, 8192
p = nb.block_params(K)
s = nb.Solver(K, T, max_in=K, max_out=N)
s.staging[:K, :T] = workload.payload(K, T, 1)
s.upload(0, K)
s.plan_encode(True, 0)
s.run()
isi = np.arange(p.Kprime, p.Kprime + N, dtype=np.uint32)
for r in range(4):
    s.mark(False)
    s.emit(isi)
    s.mark(True)
    ms = s.marked_ms()
deg = np.array([len(nb.lt_row_indices(K, int(x))) for x in isi[:512]]).mean()
print("rqb_lt_kernel: %d symbols of %d bytes in %.3f ms -> %.0f GB/s algorithmic ((deg+1)*T per symbol, deg %.2f)" % (
    N, T, ms, N * (deg + 1) * T / (ms / 1e3) / 1e9, deg))
s.close()

# a reference-format schedule replayed level by level (oracle builds it: test infrastructure, not timed)
try:
    import ctypes as C
    from oracle_lib import oracle, orc_params, ptr, u32p
    Kr = 1024
    q = orc_params(Kr)
    st = C.c_int()
    isi2 = np.arange(q.Kprime, dtype=np.uint32)
    S = oracle().orc_invert(C.byref(q), 0, ptr(isi2, u32p), C.byref(st))
    sc = S.contents
    ops = np.zeros(sc.nops, dtype=nb.api.OP_DTYPE)
    C.memmove(ops.ctypes.data, sc.ops, sc.nops * 12)
    di = np.ctypeslib.as_array(sc.di, (sc.rows,)).copy()
    c = np.ctypeslib.as_array(sc.c, (sc.cols,)).copy()
    m = nb.Matrix(sc.rows, T)
    D = np.zeros((sc.rows, T), np.uint8)
    D[q.S + q.H:q.S + q.H + Kr] = workload.payload(Kr, T, 2)
    m.upload(D)
    ms = m.schedule_replay(ops, sc.marks[0], sc.marks[1], di, c, stepwise=True)
    print("schedule replay K=%d stepwise (one row-op launch per level): %d ops, %.3f ms device" % (Kr, sc.nops, ms))
    m.upload(D)
    ms = m.schedule_replay(ops, sc.marks[0], sc.marks[1], di, c)
    print("schedule replay K=%d as one program of the solve kernel: %.3f ms device" % (Kr, ms))
    m.close()
    oracle().orc_sched_free(S)
except Exception as e:  # the capture of the LT kernel above is what matters
    print("replay skipped:", e)

# rqb_copy_rows_kernel (block image of a decoder with deferred output) and rqb_repitch_kernel
# (host<->device copies stay linear when the caller's row pitch differs from the arena's): one batch
# round trip at T = 1000 through nanorq_batch.h with page-locked buffers
Kb, Tb = 1000, 1000
Fb = Kb * Tb
pay = nb.PinnedBuffer(Fb)
pay.arr[:] = workload.payload(Kb, Tb, 3).reshape(-1)
ring = nb.PinnedBuffer((Kb + 64) * Tb)
out = nb.PinnedBuffer(Fb)
enc = nb.Encoder(Fb, Tb, Kb, 0, 8)
io_in, io_out = nb.PinnedMemIO(pay.arr), nb.PinnedMemIO(out.arr)
rows = ring.arr.reshape(Kb + 64, Tb)
assert enc.encode_range(0, 0, Kb + 64, io_in, out=rows) is not None
tags = np.array([0xFFFFFFFF if e % 10 == 3 else nb.api.tag(0, e) for e in range(Kb + 64)], np.uint32)
dec = nb.Decoder(enc.oti_common(), enc.oti_scheme_specific())
assert dec.add_symbols(tags, rows, io_out)[0] > 0 and dec.repair_block(io_out, 0)
assert np.array_equal(out.arr, pay.arr)
print("batch round trip K=%d T=%d through page-locked buffers: ok" % (Kb, Tb))
for x in (enc, dec, io_in, io_out, pay, ring, out):
    x.close()
