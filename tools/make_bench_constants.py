"""Generates nanorq_b200/bench_constants.json: the ALGORITHMIC byte counts
bench.py's roofline uses:

    W_solve = pitch * (3*N_axpy + 2*N_scal)

with N from the REFERENCE's applied op sequence (|ops| + 2*(marks[0]+1),
lib/precode.c:23-32) for the exact seeded loss patterns of nanorq_b200/workload.py,
counted by the oracle (oracle/rq_oracle.c follows the reference op for op; pinned
by tests/test_oracle.py), plus the LT-combine bytes (deg+1)*T per emitted symbol.
pitch = T rounded up to 32 (the AVX build's ALIGNED_COLS, oblas_avx.c:62).

    python tools/make_bench_constants.py        # needs oracle/liboracle.so
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nanorq_b200 import workload  # noqa: E402
from nanorq_b200.api import SolveRequest  # noqa: E402
from oracle_lib import oracle, orc_params, ptr, u32p  # noqa: E402

N_SEEDS = {"C1": 16, "C2": 16, "C3": 64, "C5": 2}


def applied(p, overhead, isi):
    st = C.c_int()
    isi = np.ascontiguousarray(isi, dtype=np.uint32)
    S = oracle().orc_invert(C.byref(p), overhead, ptr(isi, u32p), C.byref(st))
    if not S:
        return None
    na, ns = C.c_size_t(), C.c_size_t()
    n = oracle().orc_applied_ops(S, C.byref(na), C.byref(ns))
    oracle().orc_sched_free(S)
    return int(n), int(na.value), int(ns.value)


def lt_degree(p, isi):
    out = (C.c_uint32 * 40)()
    return oracle().orc_lt_indices(C.byref(p), int(isi), out)


def main():
    res = {}
    for name, (K, T, loss, oh) in workload.CONFIGS.items():
        p = orc_params(K)
        pitch = (T + 31) // 32 * 32
        n, na, ns = applied(p, 0, np.arange(p.Kprime))
        enc = {"applied_ops": n, "n_axpy": na, "n_scal": ns, "solve_bytes": pitch * (3 * na + 2 * ns)}
        dec = []
        for seed in range(N_SEEDS[name]):
            drop = workload.loss_pattern(K, loss, seed)
            extra = 0
            while True:
                esis = workload.received_esis(K, drop, oh, extra)
                req, missing = SolveRequest.for_decoder(K, esis)
                r = applied(p, req.c.overhead, req.isi)
                if r is not None:
                    break
                extra += 2
            lt = sum((lt_degree(p, e) + 1) * T for e in missing)
            dec.append({"seed": seed, "lost": int(drop.sum()), "extra": extra, "applied_ops": r[0], "n_axpy": r[1],
                        "n_scal": r[2], "solve_bytes": pitch * (3 * r[1] + 2 * r[2]), "lt_bytes": lt})
            print(name, "seed", seed, dec[-1], flush=True)
        # repair symbols an encoder emits for the same patterns: ESI K.. (ISI K'..)
        lt_rep = [(lt_degree(p, p.Kprime + k) + 1) * T for k in range(max(1024, max(d["lost"] + oh + d["extra"] for d in dec)))]
        res[name] = {"K": K, "T": T, "loss": loss, "overhead": oh, "pitch": pitch, "L": p.L, "Kprime": p.Kprime,
                     "encode": enc, "decode": dec, "lt_repair_bytes_prefix": list(np.cumsum(lt_rep).tolist()),
                     "compulsory_bytes": {"encode": K * T + p.L * T, "decode_per_lost_symbol": T}}
    with open(os.path.join(ROOT, "nanorq_b200", "bench_constants.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(k, v["encode"], "decode mean solve_bytes", np.mean([d["solve_bytes"] for d in v["decode"]]))


if __name__ == "__main__":
    main()
