#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the nanorq hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload: BASELINE.json configs[2] ("C3", the configuration the metric is quoted
on): K=4096 source symbols of T=1280 bytes per source block, 10 % Bernoulli loss,
overhead 0.  One STEP = one pass of the hot path over a batch of `--blocks`
independent source blocks per GPU: every block is encoded (intermediate-symbol
solve + the repair symbols the receiver will need) and decoded (solve on the
surviving + repair symbols, recovery of the lost symbols).
Throughput = 2 * 8 * F * blocks / t  (F = K*T payload bytes; the factor 2 because
encode and decode each process the payload -- SURVEY.md 8(d)).

  value    device-resident: symbols and solve programs already in HBM, two kernel
           launches per step (one batched solve for all encodes, one for all
           decodes), timed with CUDA events on the launching stream.
  e2e      host-to-host through the reference's own API (nanorq.h + io.h) with
           pageable host buffers: bench/rq_roundtrip.c, the SAME source that is
           compiled against the unmodified reference for `--impl reference`.
           Includes loading through ioctx, H2D, host schedule construction, the
           kernels, repair-symbol emission, decoder ingest, D2H and write-back.
  roofline the batched solve kernel against measured HBM bandwidth, algorithmic
           bytes = the reference's op sequence (constants from
           tools/make_bench_constants.py); plus the row-axpy microbenchmark
           (`row_axpy`) that genuinely streams HBM.
  cpu_baseline  the unmodified reference (oracle/_ref) on ONE host core.

`--impl reference` times the unmodified reference on all host cores on the same
workload (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, T, LOSS, OVERHEAD = 4096, 1280, 0.10, 0
F = K * T
METRIC = "encode+decode Gbit/s at K=4096, T=1280 (K'=4112), 10% loss"


# ------------------------------------------------------------------ harness
class RtConfig(C.Structure):
    _fields_ = [("K", C.c_int), ("T", C.c_int), ("nblocks", C.c_int), ("loss", C.c_double), ("overhead", C.c_int),
                ("seed", C.c_uint), ("nthreads", C.c_int), ("precalc", C.c_int), ("verify", C.c_int)]


class RtResult(C.Structure):
    _fields_ = [("wall_s", C.c_double), ("t_gen", C.c_double), ("t_emit", C.c_double), ("t_add", C.c_double),
                ("t_repair", C.c_double), ("n_lost", C.c_long), ("n_sent", C.c_long), ("retries", C.c_int),
                ("failures", C.c_int), ("mismatches", C.c_int), ("out_fnv", C.c_ulonglong)]


def roundtrip(libpath, nblocks, nthreads, seed, precalc=1, verify=1):
    L = C.CDLL(libpath)
    L.rq_roundtrip_run.argtypes = [C.POINTER(RtConfig), C.POINTER(RtResult)]
    cfg = RtConfig(K, T, nblocks, LOSS, OVERHEAD, seed, nthreads, precalc, verify)
    res = RtResult()
    rc = L.rq_roundtrip_run(C.byref(cfg), C.byref(res))
    if rc != 0 or res.failures or res.mismatches:
        raise RuntimeError("round trip failed: rc=%d failures=%d mismatches=%d (%s)" %
                           (rc, res.failures, res.mismatches, libpath))
    return res


def gbits(nblocks, seconds):
    return 2 * 8 * F * nblocks / seconds / 1e9


# ------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 50 ms while a timed region runs (recipe in
    /opt/skills/guides/B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    so = os.path.join(ROOT, "oracle", "_ref", "librq_roundtrip_ref.so")
    if not os.path.exists(so):
        raise SystemExit("oracle/_ref/librq_roundtrip_ref.so missing: run __graft_entry__.build() where /root/reference exists")
    cores = os.cpu_count() or 1
    nb = args.blocks * max(1, args.gpus)
    for w in range(args.warmup):
        roundtrip(so, min(nb, cores), cores, 100 + w, precalc=0)
    walls, parts = [], np.zeros(4)
    for s in range(args.steps):
        r = roundtrip(so, nb, cores, s, precalc=0)
        walls.append(r.wall_s)
        parts += [r.t_gen, r.t_emit, r.t_add, r.t_repair]
    t = float(np.sum(walls))
    v = gbits(nb * args.steps, t)
    sample = "%d steps x %d blocks of K=%d T=%d through nanorq.h on %d threads (bench/rq_roundtrip.c)" % (
        args.steps, nb, K, T, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C3: K=4096 T=1280 loss=10%% overhead=0, %d blocks/step" % nb, "blocks_per_step": nb},
        "cpu_baseline": {"value": v, "unit": "Gbit/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phase_seconds_summed_over_threads": dict(zip(("generate_symbols", "encode_emit", "add_symbol", "repair_block"),
                                                      [float(x) for x in parts])),
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ own arm
def build_blocks(nb_mod, nblocks, seed0):
    """Encode every block once on the GPU, derive the decoder's inputs and stage
    them; returns (encoders, decoders, checks)."""
    from nanorq_b200 import workload
    p = nb_mod.block_params(K)
    pad = p.Kprime - K
    n_rep_emit = 512  # repair symbols emitted together with the encode solve (>= lost symbols at 10 %)
    encs, decs, checks = [], [], []
    for b in range(nblocks):
        seed = seed0 + b
        src = workload.payload(K, T, seed)
        e = nb_mod.Solver(K, T, max_in=K, max_out=n_rep_emit)
        e.staging[:K, :T] = src
        e.upload(0, K)
        e.plan_encode(True, n_rep_emit)
        encs.append(e)
    nb_mod.Solver.run_batch(encs, encs[0])
    encs[0].sync()
    for b, e in enumerate(encs):
        seed = seed0 + b
        src = workload.payload(K, T, seed)
        rep = e.fetch_syms(n_rep_emit)
        drop = workload.loss_pattern(K, LOSS, seed)
        extra = 0
        while True:
            esis = workload.received_esis(K, drop, OVERHEAD, extra)
            assert len(esis) - (K - int(drop.sum())) <= n_rep_emit
            req, missing = nb_mod.SolveRequest.for_decoder(K, esis, want_c=False)  # as nanorq_repair_block does
            d = nb_mod.Solver(K, T, max_in=len(esis), max_out=len(missing))
            if d.plan(req) == 0:
                break
            d.close()
            extra += 2
        syms = np.concatenate([src[~drop], rep[:len(esis) - int((~drop).sum())]])
        d.staging[:len(esis), :T] = syms
        d.upload(0, len(esis))
        d.sync()
        decs.append(d)
        checks.append((np.asarray(missing), src[missing]))
    del pad
    return encs, decs, checks


def run_own(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import nanorq_b200 as nb
    from nanorq_b200 import sharding

    if nb.device_count() <= 0:
        raise SystemExit("no CUDA device: the nanorq_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nb.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    consts = json.load(open(os.path.join(ROOT, "nanorq_b200", "bench_constants.json")))["C3"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, device="cuda")

    NB = args.blocks
    encs, decs, checks = build_blocks(nb, NB, seed0=1000 * rank)
    own = encs[0]

    def step():
        nb.Solver.run_batch(encs, own)
        nb.Solver.run_batch(decs, own)

    # parity gate before timing: every decode returns the erased source symbols
    step()
    own.sync()
    for d, (missing, want) in zip(decs, checks):
        got = d.fetch_syms(len(missing))
        if not np.array_equal(got, want):
            raise SystemExit("decode does not reproduce the erased symbols")

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(max(args.warmup, 3)):
        step()
    own.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = nb.kernel_launches()
    own.mark(False)
    for _ in range(args.steps):
        step()
    own.mark(True)
    ms = own.marked_ms()
    barrier()
    launches = nb.kernel_launches() - l0
    ms = max_over_ranks(ms)
    ms_step = ms / args.steps
    value = gbits(NB * world, ms_step / 1e3)

    # roofline of the batched solve kernel: algorithmic bytes per launch / mean launch time
    dec_c = consts["decode"]
    enc_bytes = NB * (consts["encode"]["solve_bytes"] + consts["lt_repair_bytes_prefix"][511])
    dec_bytes = sum(dec_c[b % len(dec_c)]["solve_bytes"] + dec_c[b % len(dec_c)]["lt_bytes"] for b in range(NB))
    alg_per_launch = (enc_bytes + dec_bytes) / 2.0
    prog_bytes = 0
    for sv in encs + decs:
        st = sv.stats()
        prog_bytes += (st["n_srcs"] + st["n_gf_srcs"] + 2 * st["n_horner"] + st["n_tasks"]) * sv.pitch
    avg_launch_s = ms / 1e3 / launches
    achieved = alg_per_launch / avg_launch_s / 1e9
    compulsory = NB * (K * T + consts["L"] * T + 512 * T) + sum((K + 0) * T + len(c[0]) * T for c in checks)
    slice_bytes = nb.lib().rqb_batch_slice_bytes(NB, T)
    roofline = {"bound": "hbm", "kernel": "rqb_solve_kernel (batched, grid = %d column slices of %d bytes x %d blocks)" % (
                    -(-T // slice_bytes), slice_bytes, NB), "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_ms": 1e3 * avg_launch_s,
                "compulsory_bytes_per_step": compulsory,
                "compulsory_gbs": compulsory / (ms_step / 1e3) / 1e9,
                "program_bytes_per_launch": prog_bytes / 2.0,
                "program_gbs": prog_bytes / 2.0 / avg_launch_s / 1e9,
                "note": "achieved = ALGORITHMIC bytes of the reference's row-op sequence (3*pitch per axpy, SURVEY 8(d)) "
                        "per second. The kernel does the same linear algebra with fewer bytes: accumulations into one "
                        "destination are merged into one gather task (sources read once, destination written once) and "
                        "the H dense HDPC rows become an alpha-scan, so frac can exceed 1. program_bytes = the row "
                        "segments this kernel really loads and stores (from L1/L2/HBM); traffic = DRAM bytes per launch "
                        "from ncu (profiles/)."}
    traffic_file = os.path.join(ROOT, "profiles", "r01_solve_traffic.json")
    if os.path.exists(traffic_file):
        roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        if roofline["traffic"]:
            # what the kernel really moves through HBM (ncu, same command) per second of measured launch time
            roofline["traffic_gbs"] = roofline["traffic"] / avg_launch_s / 1e9
            roofline["traffic_frac"] = roofline["traffic_gbs"] / peak

    # ---- row-axpy microbenchmark (oaxpy, 3*T bytes per op) on a matrix >> L2
    row_axpy = None
    if not args.skip_rowaxpy:
        rows = 1 << 20
        m = nb.Matrix(rows, T)
        m.fill_random(7 + rank)
        half = rows // 2
        rng = np.random.default_rng(5)
        res = {}
        for name, beta in (("gf256", rng.integers(2, 256, half)), ("xor", np.ones(half))):
            ops = nb.Matrix.make_ops(beta, rng.permutation(half), half + rng.permutation(half))
            ol = m.upload_ops(ops)
            m.apply_dev(ol, 3)
            reps = 10
            t_ms = m.apply_dev(ol, reps)
            nb.lib().rqb_ops_free(ol)
            gbs = reps * half * 3 * m.pitch / (t_ms / 1e3) / 1e9
            res[name] = {"achieved": gbs, "frac": gbs / peak, "ms_per_launch": t_ms / reps}
        m.close()
        row_axpy = {"unit": "GB/s", "peak": peak, "ops_per_launch": half, "bytes_per_op": 3 * T,
                    "matrix_bytes": rows * T, **res}
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through nanorq.h (bench/rq_roundtrip.c)
    rt_so = os.path.join(nb.api.LIB_DIR, "librq_roundtrip.so")
    # 1.25 worker threads per host core: a thread that waits for its block's solve yields its core
    threads = max(1, min(args.threads or (5 * (os.cpu_count() or 1)) // (4 * world), 64))
    for w in range(0 if args.skip_e2e else max(args.warmup, 3)):
        roundtrip(rt_so, NB, threads, 900 + w)  # full-size steps: every worker gets its contexts and buffers
    barrier()
    nb.host_profile(reset=True)
    slow0 = nb.slow_path_counters()
    h0, d0 = nb.transfer_bytes()
    l1 = nb.kernel_launches()
    parts, t_e2e = np.zeros(4), 0.0
    for s in range(0 if args.skip_e2e else args.steps):
        # wall_s: start barrier -> last worker done, inside the harness (payload generation
        # and the byte-for-byte verification of the decoded output are outside, as in benchmark.c)
        r = roundtrip(rt_so, NB, threads, 10 * rank + s)
        parts += [r.t_gen, r.t_emit, r.t_add, r.t_repair]
        t_e2e += r.wall_s
    barrier()
    h1, d1 = nb.transfer_bytes()
    host_prof = nb.host_profile() if os.environ.get("NANORQ_B200_PROFILE") == "1" else None
    e2e_launches = nb.kernel_launches() - l1
    t_e2e = max_over_ranks(t_e2e)
    e2e = None if args.skip_e2e else {"value": gbits(NB * world * args.steps, t_e2e), "unit": "Gbit/s",
           "h2d_bytes_per_step": (h1 - h0) // args.steps, "d2h_bytes_per_step": (d1 - d0) // args.steps,
           "host_threads": threads, "api": "nanorq.h (bench/rq_roundtrip.c)", "ms_per_step": 1e3 * t_e2e / args.steps,
           "gpu_launches": e2e_launches,
           "phase_seconds_summed_over_threads": dict(zip(("generate_symbols", "encode_emit", "add_symbol", "repair_block"),
                                                         [float(x) for x in parts]))}
    if e2e is not None:
        # allocations / arena regrowths / new contexts inside the timed e2e steps: must be all zero in steady state
        e2e["slow_path_events"] = {k: v - slow0[k] for k, v in nb.slow_path_counters().items()}
    if e2e is not None and host_prof is not None:
        e2e["host_profile_seconds_summed_over_threads"] = {k: round(v, 4) for k, v in host_prof.items()}

    # ---- cpu_baseline: the unmodified reference on one host core (rank 0, N=1 only)
    cpu = None
    ref_so = os.path.join(ROOT, "oracle", "_ref", "librq_roundtrip_ref.so")
    if rank == 0 and world == 1 and not args.skip_cpu and os.path.exists(ref_so):
        roundtrip(ref_so, 2, 1, 77, precalc=0)
        tot, nblk = 0.0, 0
        while tot < args.cpu_seconds:
            r = roundtrip(ref_so, 16, 1, nblk, precalc=0)
            tot += r.wall_s
            nblk += 16
        cpu = {"value": gbits(nblk, tot), "unit": "Gbit/s", "cores": 1, "kind": "reference",
               "sample": "%d blocks of K=%d T=%d, full round trip through nanorq.h (bench/rq_roundtrip.c) on 1 of %d host "
                         "cores, unmodified reference AVX2 build (oracle/_ref)" % (nblk, K, T, os.cpu_count() or 1)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C3: K=4096 T=1280 loss=10%% overhead=0, %d independent blocks per GPU per step" % NB,
                       "blocks_per_gpu": NB, "sharding": "independent source blocks per rank, no collective",
                       "l2": "inputs larger than L2 (%.0f MB of symbols read per step)" % (2 * NB * F / 1e6),
                       "value_scope": "symbols and solve programs resident in HBM; host schedule construction is inside e2e"},
            "roofline": roofline, "row_axpy": row_axpy, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
        }))
    for s in encs + decs:
        s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--blocks", type=int, default=118,
                    help="source blocks per GPU per step (118 blocks x 5 column slices of 256 bytes fill the 148 x 4 resident CTA slots once)")
    ap.add_argument("--threads", type=int, default=0, help="host threads for the e2e arm (default: cores / ranks)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-rowaxpy", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
