#!/usr/bin/env python
"""Solve-kernel figures for every BASELINE.json configuration (not a bench line: a table for DESIGN.md).
Per configuration: device time of one block on its own (the flavour the planner picks) and the throughput of
one batched launch over many blocks, encode and decode, CUDA events on the launching stream, programs
built before the clock.  Every decoded block is checked against its payload once.

    python tools/kernel_table.py            (on a GPU box)  -> gpurun_out/kernel_table.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nanorq_b200 as nb  # noqa: E402
from nanorq_b200 import workload  # noqa: E402

# name -> blocks per batched launch
BATCH = {"C1": 255, "C2": 236, "C3": 118, "C5": 32}


def median_ms(fn, reps):
    out = []
    for _ in range(reps):
        out.append(fn())
    return float(np.median(out[1:])) if len(out) > 1 else out[0]


def main():
    rows = []
    for name, (K, T, loss, oh) in workload.CONFIGS.items():
        p = nb.block_params(K)
        nblk = BATCH[name]
        loss = loss if loss > 0 else 0.3  # C1 has no loss: a decode needs something to recover
        src = [workload.payload(K, T, b) for b in range(min(nblk, 16))]
        encs, decs = [], []
        n_rep_max = int(K * loss * 2) + oh + 8
        for b in range(nblk):
            e = nb.Solver(K, T, max_in=K, max_out=n_rep_max, flavour="hbm" if nblk > 1 else "auto")
            e.staging[:K, :T] = src[b % len(src)]
            e.upload(0, K)
            e.plan_encode(True, 0)
            encs.append(e)
        nb.Solver.run_batch(encs, encs[0]) if nblk > 1 else encs[0].run()
        encs[0].sync()
        # repair symbols of every block, then the decoders with their own loss pattern
        checks = []
        for b in range(nblk):
            drop = workload.loss_pattern(K, loss, b)
            esis = workload.received_esis(K, drop, oh)
            need = len(esis) - int((~drop).sum())
            encs[b].emit(np.arange(K, K + need, dtype=np.uint32) + (p.Kprime - K))
            rep = encs[b].fetch_syms(need)
            while True:
                req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
                d = nb.Solver(K, T, max_in=len(esis) + 8, max_out=len(missing), flavour="hbm" if nblk > 1 else "auto")
                d.staging[:len(esis), :T] = np.concatenate([src[b % len(src)][~drop], rep])
                d.upload(0, len(esis))
                if d.plan(req) == 0:
                    break
                d.close()  # singular: two more repair symbols, as a receiver would ask for
                encs[b].emit(np.arange(K + need, K + need + 2, dtype=np.uint32) + (p.Kprime - K))
                rep = np.concatenate([rep, encs[b].fetch_syms(2)])
                esis = np.concatenate([esis, np.arange(K + need, K + need + 2, dtype=np.uint32)])
                need += 2
            decs.append(d)
            checks.append((missing, src[b % len(src)]))
        nb.Solver.run_batch(decs, decs[0]) if nblk > 1 else decs[0].run()
        for d, (missing, s) in zip(decs, checks):
            assert np.array_equal(d.fetch_syms(len(missing)), s[missing]), name

        def batch_ms(group):
            own = group[0]
            own.mark(False)
            nb.Solver.run_batch(group, own) if len(group) > 1 else own.run()
            own.mark(True)
            return own.marked_ms()

        def single_ms(s):
            s.mark(False)
            s.run()
            s.mark(True)
            return s.marked_ms()

        bits = 8.0 * K * T
        # one block on its own, in the flavour the planner picks for it
        e_one = nb.Solver(K, T, max_in=K, max_out=16)
        e_one.staging[:K, :T] = src[0]
        e_one.upload(0, K)
        e_one.plan_encode(True, 0)
        drop = workload.loss_pattern(K, loss, 0)
        esis = workload.received_esis(K, drop, oh + 4)
        req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
        d_one = nb.Solver(K, T, max_in=len(esis), max_out=len(missing))
        d_one.staging[:len(esis), :T] = 7
        d_one.upload(0, len(esis))
        assert d_one.plan(req) == 0
        e1, d1 = median_ms(lambda: single_ms(e_one), 8), median_ms(lambda: single_ms(d_one), 8)
        st_e, st_d = e_one.stats(), d_one.stats()
        e_one.close()
        d_one.close()
        eb, db = median_ms(lambda: batch_ms(encs), 6), median_ms(lambda: batch_ms(decs), 6)
        rows.append({"config": name, "K": K, "T": T, "loss": loss, "overhead": oh,
                     "flavour_single": "smem" if st_e.get("smem") else "hbm",
                     "single_encode_ms": round(e1, 4), "single_decode_ms": round(d1, 4),
                     "levels_encode": st_e["n_levels"], "levels_decode": st_d["n_levels"],
                     "blocks_per_launch": nblk, "batch_encode_ms": round(eb, 4), "batch_decode_ms": round(db, 4),
                     "batch_encode_gbit_s": round(nblk * bits / eb / 1e6, 1),
                     "batch_decode_gbit_s": round(nblk * bits / db / 1e6, 1),
                     "batch_encode_decode_gbit_s": round(2 * nblk * bits / (eb + db) / 1e6, 1)})
        print(json.dumps(rows[-1]), flush=True)
        for s in encs + decs:
            s.close()
        nb.release_cached()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"rows": rows}, open(os.path.join(ROOT, "gpurun_out", "kernel_table.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
