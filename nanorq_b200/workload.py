"""Seeded synthetic workloads shared by bench.py, the tests and
tools/make_bench_constants.py (so that the roofline constants are computed for
exactly the loss patterns the benchmark runs).  Host-side integer bookkeeping
only -- no symbol arithmetic."""
import numpy as np

# BASELINE.json configs: name -> (K, T, loss, overhead)
CONFIGS = {
    "C1": (10, 64, 0.0, 0),
    "C2": (1024, 1280, 0.05, 2),
    "C3": (4096, 1280, 0.10, 0),
    "C5": (56403, 512, 0.15, 0),
}


def payload(K, T, seed):
    return np.random.default_rng(1000003 * seed + 17).integers(0, 256, (K, T), dtype=np.uint8)


def loss_pattern(K, loss, seed):
    """Bernoulli(loss) drop mask over the K source symbols (at least one symbol is
    dropped when loss > 0 so that a decode always has work to do)."""
    drop = np.random.default_rng(7919 * seed + 3).random(K) < loss
    if loss > 0 and not drop.any():
        drop[seed % K] = True
    return drop


def received_esis(K, drop, overhead, extra=0):
    """Surviving source ESIs in order, then repair ESIs K, K+1, ... : one per
    dropped symbol plus `overhead` (+ `extra` after a singular verdict)."""
    n_rep = int(drop.sum()) + overhead + extra
    return np.concatenate([np.nonzero(~drop)[0], np.arange(K, K + n_rep)]).astype(np.uint32)
