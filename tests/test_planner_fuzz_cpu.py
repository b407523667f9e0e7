"""CPU: seeded random requests through the planner and the program interpreter (oracle/plan_interp.c),
against the oracle.  Random block sizes, loss, overhead, arrival order and -- for the shared-memory
flavour -- random slot budgets, so that every table width (4..8 bits), the scan chunking and the
fall-back to the HBM flavour are exercised, not only the kernel's own budget."""
import numpy as np
import pytest

import nanorq_b200 as nb
from nanorq_b200 import api
from oracle_lib import interp_run, orc_decode, orc_encode, orc_lt, orc_params


def one_request(seed):
    rng = np.random.default_rng(seed)
    K = int(rng.choice([rng.integers(1, 30), rng.integers(30, 300), rng.integers(300, 1500), rng.integers(300, 1500), rng.integers(1500, 4200)]))
    T = int(rng.choice([8, 16, 24]))
    p = orc_params(K)
    src = rng.integers(0, 256, (K, T), dtype=np.uint8)
    Cm, _, _ = orc_encode(K, T, src)
    full = api.smem_budget()
    budget = int(rng.choice([0, full, rng.integers(full // 64, full), rng.integers(full // 8, full)]))
    if rng.random() < 0.3:  # an encoder request: all intermediate symbols + a window of repair symbols
        out_isi = np.arange(K, K + int(rng.integers(1, 40)), dtype=np.uint32) + (p.Kprime - K)
        req = nb.SolveRequest.for_encoder(K, True, out_isi)
        rc, blob = nb.plan_blob(K, req, smem=budget)
        assert rc == 0, (seed, rc)
        rc, cout, sout = interp_run(blob, src, T, p.L, len(out_isi))
        assert rc == 0, (seed, rc)
        assert np.array_equal(cout, Cm), seed
        assert np.array_equal(sout, np.stack([orc_lt(K, T, Cm, int(x)) for x in out_isi])), seed
        return blob
    loss = float(rng.choice([0.02, 0.1, 0.3, 0.6, 0.95]))
    oh = int(rng.choice([0, 0, 1, 2, 5, 20]))
    drop = rng.random(K) < loss
    if not drop.any():
        drop[int(rng.integers(0, K))] = True
    rep = K + rng.choice(4 * K + 64, size=int(drop.sum()) + oh, replace=False)  # repair ESIs anywhere, not a run
    esis = np.concatenate([np.nonzero(~drop)[0], rep]).astype(np.uint32)
    rng.shuffle(esis)
    syms = np.stack([src[e] if e < K else orc_lt(K, T, Cm, int(e) + p.Kprime - K) for e in esis])
    rc_o, out_o, C_o, _, _ = orc_decode(K, T, esis, syms, want_C=True)
    req, missing = nb.SolveRequest.for_decoder(K, esis)
    rc_p, blob = nb.plan_blob(K, req, smem=budget)
    assert (rc_o == 0) == (rc_p == 0), (seed, K, rc_o, rc_p)
    if rc_o != 0:
        return None
    rc, cout, sout = interp_run(blob, syms, T, p.L, len(missing))
    assert rc == 0, (seed, rc)
    assert np.array_equal(sout, src[missing]), seed
    return blob


@pytest.mark.parametrize("chunk", range(6))
def test_random_requests_on_the_interpreter(chunk):
    seen = set()
    for seed in range(chunk * 100, chunk * 100 + 100):
        blob = one_request(31000 + seed)
        if blob is not None:
            seen.add((blob["smem"], blob["slice_bytes"], blob["tab_bits"]))
    print(sorted(seen))
