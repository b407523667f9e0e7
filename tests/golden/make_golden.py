#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference): it compiles the reference
through oracle/Makefile (-> oracle/_ref/libnanorq_ref.so) and records
  * kat.json   -- FNV-1a-64 hashes of source symbols, 16 repair symbols and all
                  intermediate symbols for the SURVEY section 8(c) configurations,
                  plus schedule sizes (|ops|, marks, i, u);
  * small.npz  -- full byte vectors for small blocks: intermediate symbols,
                  repair symbols and one decode case each (ESIs fed + recovered bytes).
The GPU box has no /root/reference; tests there compare against these files."""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import fnv1a64, kat_payload, orc_params, ptr, ref, u32p  # noqa: E402


def ref_encode(K, T, src, want_ops=False):
    R = ref()
    p = orc_params(K)
    D = np.zeros((p.L, T), np.uint8)
    D[p.S + p.H:p.S + p.H + K] = src.reshape(K, T)
    isi = np.arange(p.Kprime, dtype=np.uint32)
    Cm = np.zeros((p.L, T), np.uint8)
    info = (C.c_long * 5)()
    rc = R.ref_solve(K, T, 0, ptr(isi, u32p), ptr(D), ptr(Cm), C.byref(info), None, 0)
    assert rc == 0
    return Cm, list(info)


def ref_lt(K, T, Cm, isi):
    out = np.zeros(T, np.uint8)
    ref().ref_lt_row(K, T, ptr(Cm), isi, ptr(out))
    return out


def main():
    kat = {}
    for K, T in ((10, 64), (1024, 1280), (4096, 1280), (56403, 512)):
        p = orc_params(K)
        src = kat_payload(K * T)
        Cm, info = ref_encode(K, T, src)
        rep = np.concatenate([ref_lt(K, T, Cm, e + p.Kprime - K) for e in range(K, K + 16)])
        kat["%d,%d" % (K, T)] = {
            "fnv_source": "%016x" % fnv1a64(src), "fnv_repair16": "%016x" % fnv1a64(rep),
            "fnv_intermediate": "%016x" % fnv1a64(Cm), "nops": info[0], "marks": info[1:3],
            "i": info[3], "u": info[4]}
        print(K, T, kat["%d,%d" % (K, T)])
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1, sort_keys=True)

    small = {}
    rng = np.random.default_rng(20260101)
    for K, T, loss, oh in ((10, 64, 0.3, 0), (26, 16, 0.4, 1), (101, 24, 0.2, 0), (257, 8, 0.15, 3)):
        p = orc_params(K)
        src = rng.integers(0, 256, K * T, dtype=np.uint8)
        Cm, _ = ref_encode(K, T, src)
        drop = rng.random(K) < loss
        keep = np.nonzero(~drop)[0]
        esis = np.concatenate([keep, np.arange(K, K + int(drop.sum()) + oh)]).astype(np.uint32)
        rng.shuffle(esis)
        syms = np.zeros((len(esis), T), np.uint8)
        oti = (C.c_uint64 * 2)()
        rc = ref().ref_encode_api(K, T, ptr(src), ptr(esis, u32p), len(esis), ptr(syms), C.byref(oti), None, None, 0)
        assert rc == 0
        out = np.zeros(K * T, np.uint8)
        rc = ref().ref_decode_api(C.byref(oti), T, ptr(esis, u32p), ptr(syms), len(esis), ptr(out), K * T, None, None)
        key = "K%d_T%d" % (K, T)
        small[key + "_src"] = src
        small[key + "_C"] = Cm
        small[key + "_esis"] = esis
        small[key + "_syms"] = syms
        small[key + "_rc"] = np.array([rc])
        small[key + "_out"] = out
        small[key + "_oti"] = np.array([oti[0], oti[1]], dtype=np.uint64)
        print(key, "decode rc", rc, "roundtrip", np.array_equal(out, src))
    np.savez_compressed(os.path.join(HERE, "small.npz"), **small)


if __name__ == "__main__":
    main()
