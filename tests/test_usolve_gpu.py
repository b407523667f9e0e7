"""The Schur elimination on the device (rqb_usolve_kernel, warp-level pivot search) against the host
code of the planner: same pivots, so the emitted programs must be byte-identical; same verdict on
singular systems."""
import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import workload
from oracle_lib import orc_decode, orc_encode, orc_lt, orc_params

pytestmark = pytest.mark.gpu


def plans(K, req, smem):
    out = []
    for mode in (1, 2):  # host, device
        nb.lib().rqb_set_usolve_mode(mode)
        out.append(nb.plan_blob(K, req, smem=smem))
    nb.lib().rqb_set_usolve_mode(0)
    return out


@pytest.mark.parametrize("K,loss,oh,smem", [(10, 0.4, 0, True), (100, 0.3, 1, True), (1024, 0.05, 2, True),
                                             (4096, 0.10, 0, True), (4096, 0.10, 0, False), (4096, 0.3, 40, False),
                                             (20000, 0.2, 0, False), (56403, 0.15, 0, False)])
def test_device_usolve_gives_the_same_program_as_the_host(K, loss, oh, smem):
    singular = 0
    for seed in range(3 if K < 10000 else 1):
        drop = workload.loss_pattern(K, loss, seed)
        esis = workload.received_esis(K, drop, oh, 0)
        req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
        (rc_h, blob_h), (rc_d, blob_d) = plans(K, req, smem)
        assert rc_h == rc_d, (K, seed, rc_h, rc_d)
        if rc_h != 0:
            singular += 1
            continue
        assert blob_h["n_pages"] == blob_d["n_pages"]
        assert np.array_equal(blob_h["pages"], blob_d["pages"]), (K, seed)
        for k in ("i", "u", "rho", "nfree", "n_tasks", "n_levels"):
            assert blob_h["stats"][k] == blob_d["stats"][k], k
    print("K=%d: %d singular patterns (verdicts agree)" % (K, singular))


@pytest.mark.parametrize("K,loss", [(10, 0.5), (100, 0.3), (1024, 0.1), (4096, 0.1)])
def test_device_usolve_reports_singular_systems_like_the_host(K, loss):
    """A system made singular on purpose (the same repair symbol delivered for two missing source
    symbols) and many random patterns at overhead 0 (about one in a hundred is singular at K=10):
    the verdict of the device elimination is the host's."""
    smem = K <= 4096
    drop = workload.loss_pattern(K, loss, 1)
    if drop.sum() < 2:
        drop[:2] = True
    esis = workload.received_esis(K, drop, 0, 0)
    req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
    isi = req.isi.copy()
    isi[missing[1]] = isi[missing[0]]  # two rows of the constraint matrix are now equal
    bad = nb.SolveRequest(isi, req.in_row, req.c.overhead, False, missing)
    (rc_h, _), (rc_d, _) = plans(K, bad, smem)
    assert rc_h == 1 and rc_d == 1, (rc_h, rc_d)  # RQB_NEED_MORE from both
    seen = {0: 0, 1: 0}
    for seed in range(600 if K == 10 else 12):
        rng = np.random.default_rng(seed)
        d2 = rng.random(K) < loss
        if not d2.any():
            continue
        e2 = workload.received_esis(K, d2, 0, 0)
        r2, _ = nb.SolveRequest.for_decoder(K, e2, want_c=False)
        (rc_h, _), (rc_d, _) = plans(K, r2, smem)
        assert rc_h == rc_d, seed
        seen[1 if rc_h else 0] += 1
    print("K=%d decodable / singular random patterns:" % K, seen)


def test_c5_round_trip_with_the_usolve_on_the_device():
    """K=56403: the default mode runs the elimination on the device (u > 256); decode vs the source."""
    K, T = 56403, 64
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (K, T), dtype=np.uint8)
    p = nb.block_params(K)
    e = nb.Solver(K, T, max_in=K, max_out=10000)
    e.staging[:K, :T] = src
    e.upload(0, K)
    e.plan_encode(True, 0)
    e.run()
    drop = workload.loss_pattern(K, 0.15, 2)
    nrep = int(drop.sum()) + 4
    e.emit(np.arange(p.Kprime, p.Kprime + nrep, dtype=np.uint32))
    rep = e.fetch_syms(nrep)
    e.close()
    esis = workload.received_esis(K, drop, 4, 0)
    req, missing = nb.SolveRequest.for_decoder(K, esis, want_c=False)
    d = nb.Solver(K, T, max_in=len(esis), max_out=len(missing))
    d.staging[:len(esis), :T] = np.concatenate([src[~drop], rep])
    d.upload(0, len(esis))
    l0 = nb.kernel_launches()
    assert d.plan(req) == 0
    assert nb.kernel_launches() - l0 == 1  # the usolve kernel ran inside the planning
    d.run()
    assert np.array_equal(d.fetch_syms(len(missing)), src[missing])
    d.close()
