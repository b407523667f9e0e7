"""K' table sweep (include/table2.h:6-51 in the reference, 477 rows): for every K'
and for K'-1 (one padding symbol) the product planner's encode program and one
20 %-loss decode program are run on the CPU interpreter and compared with the
oracle -- intermediate symbols, recovered symbols and the decode verdict.
The same sweep on a subset runs on the GPU (-m gpu) through the solver C-ABI.

CPU budget: every K' <= 10000 plus every 6th larger one by default (about a minute);
NANORQ_FULL_SWEEP=1 runs all 477 rows (about four minutes)."""
import os
import re

import numpy as np
import pytest

import nanorq_b200 as nb
from oracle_lib import interp_run, kat_payload, orc_decode, orc_encode, orc_lt, orc_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TXT = open(os.path.join(ROOT, "nanorq_b200", "csrc", "rfc6330_tables.h")).read()
KPRIMES = [int(m) for m in re.findall(r"\{\s*(\d+),\s*\d+,\s*\d+,\s*\d+,\s*\d+\s*\}", _TXT)]
FULL = os.environ.get("NANORQ_FULL_SWEEP") == "1"


def test_table_has_477_rows():
    assert len(KPRIMES) == 477 and KPRIMES[0] == 10 and KPRIMES[-1] == 56403


def sweep_values():
    big = [k for k in KPRIMES if k > 10000]
    keep = set(k for k in KPRIMES if k <= 10000) | set(big if FULL else big[::6]) | {56403}
    return sorted(keep)


def one_K(K, T, run_encode, run_decode):
    """encode + one decode of K source symbols, compared with the oracle; returns
    the number of singular decodes (verdicts must agree)."""
    p = orc_params(K)
    src = kat_payload(K * T).reshape(K, T)
    Cm, _, _ = orc_encode(K, T, src)
    out_isi = np.arange(p.Kprime, p.Kprime + 6, dtype=np.uint32)
    cout, sout = run_encode(K, T, src, out_isi, p)
    assert np.array_equal(cout, Cm), K
    rep = np.stack([orc_lt(K, T, Cm, int(x)) for x in out_isi])
    assert np.array_equal(sout, rep), K
    rng = np.random.default_rng(K)
    drop = rng.random(K) < 0.2
    if not drop.any():
        drop[int(rng.integers(0, K))] = True
    nd = int(drop.sum())
    esis = np.concatenate([np.nonzero(~drop)[0], np.arange(K, K + nd + 1)]).astype(np.uint32)  # overhead 1
    syms = np.stack([src[e] if e < K else orc_lt(K, T, Cm, int(e) + p.Kprime - K) for e in esis])
    rc_o, out_o, _, _, _ = orc_decode(K, T, esis, syms)
    rc_p, got = run_decode(K, T, esis, syms, p)
    assert (rc_o == 0) == (rc_p == 0), (K, rc_o, rc_p)
    if rc_o != 0:
        return 1
    assert np.array_equal(got, src[np.nonzero(drop)[0]]), K
    return 0


SMEM = True  # ask for the shared-memory flavour: the planner falls back to the HBM one when the rows do not fit


def interp_encode(K, T, src, out_isi, p):
    rc, blob = nb.plan_blob(K, nb.SolveRequest.for_encoder(K, True, out_isi), smem=SMEM)
    assert rc == 0
    FLAVOURS[bool(blob["smem"])] = FLAVOURS.get(bool(blob["smem"]), 0) + 1
    rc, cout, sout = interp_run(blob, src, T, p.L, len(out_isi))
    assert rc == 0, (K, rc)
    return cout, sout


def interp_decode(K, T, esis, syms, p):
    req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
    rc, blob = nb.plan_blob(K, req, smem=SMEM)
    if rc != 0:
        return rc, None
    rc, _, sout = interp_run(blob, syms, T, 0, len(missing))
    assert rc == 0, (K, rc)
    return 0, sout


FLAVOURS = {}


def test_every_kprime_on_the_interpreter():
    global SMEM
    singular = 0
    vals = sweep_values()
    for Kp in vals:
        SMEM = True
        singular += one_K(Kp, 8, interp_encode, interp_decode)
        if Kp <= 3000 or Kp == 56403:  # K'-1: one padding symbol, same K'
            singular += one_K(Kp - 1, 8, interp_encode, interp_decode)
        if Kp <= 3000:  # the HBM flavour as well where the shared-memory one is the default
            SMEM = False
            singular += one_K(Kp, 8, interp_encode, interp_decode)
    print("%d K' values, %d singular decodes (verdicts agree); encode programs by flavour: %s" % (
        len(vals), singular, {("smem" if k else "hbm"): v for k, v in FLAVOURS.items()}))
    assert FLAVOURS.get(True, 0) > 100 and FLAVOURS.get(False, 0) > 50


# ------------------------------------------------------------------ GPU subset
def gpu_encode_run(K, T, src, out_isi, p):
    s = nb.Solver(K, T, max_in=K, max_out=len(out_isi))
    s.staging[:K, :T] = src
    s.upload(0, K)
    assert s.plan(nb.SolveRequest.for_encoder(K, True, out_isi)) == 0
    s.run()
    cout, sout = s.fetch_c(), s.fetch_syms(len(out_isi))
    s.close()
    return cout, sout


def gpu_decode_run(K, T, esis, syms, p):
    req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
    d = nb.Solver(K, T, max_in=len(esis), max_out=len(missing))
    d.staging[:len(esis), :T] = syms
    d.upload(0, len(esis))
    rc = d.plan(req)
    got = None
    if rc == 0:
        d.run()
        got = d.fetch_syms(len(missing))
    d.close()
    return rc, got


@pytest.mark.gpu
def test_kprime_subset_on_the_gpu():
    """30 rows of the K' table spread over its whole range (and K'-1 for the small ones)
    through rqb_solver_* on the device, against the oracle."""
    step = len(KPRIMES) // 29
    vals = sorted(set(KPRIMES[::step]) | {10, 56403})
    singular = 0
    for Kp in vals:
        singular += one_K(Kp, 48, gpu_encode_run, gpu_decode_run)
        if Kp <= 3000:
            singular += one_K(Kp - 1, 48, gpu_encode_run, gpu_decode_run)
    print("%d K' values on the GPU, %d singular decodes" % (len(vals), singular))
