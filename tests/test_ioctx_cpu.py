"""ioctx adapters (include/io.h) against the reference's lib/io.c: the same operation
sequence through both libraries' function tables must give the same results and files.
CPU only: the adapters are host code."""
import ctypes as C
import os

import numpy as np
import pytest

import nanorq_b200 as nb
from oracle_lib import have_ref, ref

SZ = C.c_size_t


class IoCtx(C.Structure):
    pass


IoCtx._fields_ = [("read", C.CFUNCTYPE(SZ, C.POINTER(IoCtx), C.c_void_p, SZ)),
                  ("write", C.CFUNCTYPE(SZ, C.POINTER(IoCtx), C.c_void_p, SZ)),
                  ("seek", C.CFUNCTYPE(C.c_bool, C.POINTER(IoCtx), SZ)),
                  ("size", C.CFUNCTYPE(SZ, C.POINTER(IoCtx))),
                  ("tell", C.CFUNCTYPE(C.c_long, C.POINTER(IoCtx))),
                  ("destroy", C.CFUNCTYPE(None, C.POINTER(IoCtx))),
                  ("seekable", C.c_bool), ("writable", C.c_bool)]


def libs():
    out = [("b200", nb.lib())]
    if have_ref():
        out.append(("reference", ref()))
    for _, L in out:
        for f in ("ioctx_from_file", "ioctx_mmap_file"):
            getattr(L, f).restype = C.POINTER(IoCtx)
            getattr(L, f).argtypes = [C.c_char_p, C.c_int]
        L.ioctx_from_mem.restype = C.POINTER(IoCtx)
        L.ioctx_from_mem.argtypes = [C.c_void_p, SZ]
    return out


def script(io, log):
    """a decoder-style then encoder-style use of one ioctx; appends every result to log"""
    c = io.contents
    buf = (C.c_uint8 * 4096)()
    data = np.arange(4096, dtype=np.uint32).astype(np.uint8)
    log.append(("flags", c.seekable, c.writable))
    for off, n in [(0, 100), (1000, 1280), (100, 900), (5000, 77), (2280, 1)]:
        ok = c.seek(io, off)
        w = c.write(io, data[off % 256:].ctypes.data, n) if ok else -1
        log.append(("write", off, n, bool(ok), int(w), int(c.tell(io)) if ok else None))
    log.append(("size", int(c.size(io))))
    for off, n in [(0, 50), (990, 20), (2279, 10), (5070, 100)]:
        ok = c.seek(io, off)
        r = c.read(io, buf, n) if ok else -1
        log.append(("read", off, n, bool(ok), int(r), bytes(buf[:max(r, 0)])))


@pytest.mark.parametrize("maker", ["ioctx_from_file", "ioctx_mmap_file"])
def test_file_backed_ioctx_matches_reference(tmp_path, maker):
    results = {}
    for name, L in libs():
        path = str(tmp_path / ("%s_%s.bin" % (maker, name))).encode()
        io = getattr(L, maker)(path, 0)  # create (decoder side)
        assert bool(io), name
        log = []
        script(io, log)
        io.contents.destroy(io)
        content = open(path, "rb").read()
        # reopen for reading (encoder side)
        io = getattr(L, maker)(path, 1)
        assert bool(io), name
        c = io.contents
        buf = (C.c_uint8 * 2000)()
        log.append(("ro_flags", c.seekable, c.writable, int(c.size(io))))
        for off, n in [(0, 1280), (1000, 1280), (len(content) - 10, 100)]:
            ok = c.seek(io, off)
            r = c.read(io, buf, n) if ok else -1
            log.append(("ro_read", off, bool(ok), int(r), bytes(buf[:max(r, 0)])))
        c.destroy(io)
        results[name] = (log, content)
    log, content = results["b200"]
    assert content[1000:2280] == bytes(np.arange(4096, dtype=np.uint32).astype(np.uint8)[1000 % 256:][:1280])
    assert len(content) == 5077
    if "reference" in results:
        rlog, rcontent = results["reference"]
        assert content == rcontent
        for a, b in zip(log, rlog):
            if maker == "ioctx_mmap_file" and a[0] == "size":
                # the reference reports the physical size of its file, which it grows in 64 KiB
                # windows while writing (lib/io.c:323-331); this build reports the bytes written.
                # The file itself is identical after destroy() (checked above).
                assert a[1] == 5077 and b[1] % 65536 == 0
                continue
            assert a == b, (a, b)
    assert not getattr(nb.lib(), maker)(str(tmp_path / "missing" / "x").encode(), 1)  # NULL for a missing file


def test_mem_ioctx_matches_reference():
    results = {}
    for name, L in libs():
        arr = np.zeros(3000, dtype=np.uint8)
        io = L.ioctx_from_mem(arr.ctypes.data, arr.size)
        c = io.contents
        log = [("flags", c.seekable, c.writable, int(c.size(io)))]
        data = (np.arange(3000) % 251).astype(np.uint8)
        buf = (C.c_uint8 * 4000)()
        for off, n in [(0, 1280), (1280, 1280), (2560, 1280), (2999, 5), (3000, 1), (10, 0)]:
            ok = c.seek(io, off)
            w = c.write(io, data[off % 100:].ctypes.data, n) if ok else -1
            log.append(("write", off, n, bool(ok), int(w), int(c.tell(io))))
        for off, n in [(0, 3000), (2990, 100), (3000, 10)]:
            ok = c.seek(io, off)
            r = c.read(io, buf, n) if ok else -1
            log.append(("read", off, n, bool(ok), int(r), bytes(buf[:max(r, 0)])))
        c.destroy(io)
        results[name] = (log, arr.copy())
    if "reference" in results:
        assert np.array_equal(results["b200"][1], results["reference"][1])
        for a, b in zip(results["b200"][0], results["reference"][0]):
            assert a == b, (a, b)
    assert results["b200"][1][2999] != 0
