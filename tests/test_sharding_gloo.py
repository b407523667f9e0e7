"""N>1 path on CPU: world_size-2 gloo processes shard the source blocks of one
object (SBN b -> rank b % world), each rank plans ITS blocks with the product's
host planner and -- because there is no GPU here -- executes the device program
on the test-side interpreter; the per-block digests are gathered and compared
with the oracle.  Also checks the partition and the max-over-ranks reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F_T, F_K, F_Z = 32, 96, 5  # 5 source blocks of 96 symbols of 32 bytes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nanorq_b200 as nb
    from nanorq_b200 import sharding
    from oracle_lib import fnv1a64, interp_run, kat_payload, orc_params

    F = F_T * F_K * F_Z
    enc = nb.Encoder(F, F_T, F_K, 0, 8)  # host-side object: partitioning only
    Z = enc.blocks()
    payload = kat_payload(F).reshape(-1, F_T)
    mine = sharding.blocks_for_rank(Z, rank, world)
    digests = {}
    first = 0
    starts = []
    for sbn in range(Z):
        starts.append(first)
        first += enc.block_symbols(sbn)
    for sbn in mine:
        K = enc.block_symbols(sbn)
        src = payload[starts[sbn]:starts[sbn] + K]
        p = orc_params(K)
        rc, blob = nb.plan_blob(K, nb.SolveRequest.for_encoder(K, True, ()))
        assert rc == 0
        rc, cout, _ = interp_run(blob, src, F_T, p.L, 0)
        assert rc == 0
        digests[sbn] = fnv1a64(cout)
    everyone = sharding.gather_objects((rank, mine, digests))
    slowest = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((Z, everyone, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_blocks_for_rank_is_a_partition():
    from nanorq_b200 import sharding
    for n, world in [(8, 1), (8, 2), (8, 8), (5, 2), (3, 8), (256, 7)]:
        seen = []
        for r in range(world):
            mine = sharding.blocks_for_rank(n, r, world)
            assert all(sharding.owner_of(b, world) == r for b in mine)
            seen += mine
        assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        sharding.blocks_for_rank(4, 2, 2)


def test_two_gloo_ranks_shard_blocks_and_match_oracle():
    from oracle_lib import fnv1a64, kat_payload, orc_encode
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    Z, everyone, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert Z == F_Z and slowest == 11.0  # max over ranks of 10 + rank
    owned = sorted(b for _, mine, _ in everyone for b in mine)
    assert owned == list(range(Z))  # every block exactly once, no overlap
    payload = kat_payload(F_T * F_K * F_Z).reshape(-1, F_T)
    got = {}
    for _, _, d in everyone:
        got.update(d)
    for sbn in range(Z):
        Cm, _, _ = orc_encode(F_K, F_T, payload[sbn * F_K:(sbn + 1) * F_K])
        assert got[sbn] == fnv1a64(Cm), "block %d solved on its rank differs from the oracle" % sbn
