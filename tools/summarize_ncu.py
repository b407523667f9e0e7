#!/usr/bin/env python
"""Turn the ncu captures that tools/profile_solve.sh left in gpurun_out/ into the
committed summaries under profiles/ (run in the build container; needs only `ncu -i`).

    python tools/summarize_ncu.py r01
writes profiles/<round>_launches.csv          the per-launch device times (ncu launch list)
       profiles/<round>_launch_shares.json    share of the step per kernel
       profiles/<round>_solve_ncu.json        key metrics of rqb_solve_kernel (per launch)
       profiles/<round>_rowops_ncu.json       key metrics of rqb_rowops_kernel
       profiles/<round>_aux_ncu.json          key metrics of rqb_lt_kernel / rqb_gather_rows_kernel (tools/aux_kernels.py)
       profiles/<round>_solve_traffic.json    DRAM bytes per launch (bench.py's roofline.traffic)
"""
import csv
import hashlib
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEYS or h == "Kernel Name":
                try:
                    d[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    d[h] = v
        out.append(d)
    return out


def source_stamp():
    """Same stamp as bench.py's traffic_stamp(): the sources that decide what the solve kernel moves."""
    h = hashlib.sha256()
    for f in ("nanorq_b200/csrc/rqb_device.cu", "nanorq_b200/csrc/rqb_planner.c", "nanorq_b200/csrc/rqb_program.h"):
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(m):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return m["value"] * scale[m["unit"]]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    # launch list
    src = os.path.join(OUT, "launches.csv")
    if os.path.exists(src):
        lines = [l for l in open(src) if l.startswith('"')]
        with open(os.path.join(PROF, rnd + "_launches.csv"), "w") as f:
            f.writelines(lines)
        tot = {}
        for r in csv.DictReader(io.StringIO("".join(lines))):
            k = r["Kernel Name"].split("(")[0]
            t = tot.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += float(r["Metric Value"].replace(",", "")) / 1e6
        total = sum(v[1] for v in tot.values())
        shares = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / total, 4)} for k, v in tot.items()}
        json.dump({"command": "python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e (under ncu --metrics gpu__time_duration.sum --clock-control none)",
                   "note": "per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes",
                   "kernels": shares}, open(os.path.join(PROF, rnd + "_launch_shares.json"), "w"), indent=1)
        print(json.dumps(shares, indent=1))
    for name in ("solve", "smem", "rowops", "aux"):
        rep = os.path.join(OUT, "prof_%s.ncu-rep" % name)
        if not os.path.exists(rep):
            continue
        rows = raw(rep)
        json.dump({"capture": "ncu --set full --clock-control none --import-source on -k regex:rqb_%s (bench.py kernel arm)" % name,
                   "launches": rows}, open(os.path.join(PROF, "%s_%s_ncu.json" % (rnd, name)), "w"), indent=1)
        if name == "solve":
            per = [to_bytes(r["dram__bytes_read.sum"]) + to_bytes(r["dram__bytes_write.sum"]) for r in rows]
            grid = rows[0].get("launch__grid_size")
            slices = 5  # T = 1280 in 256-byte column slices
            json.dump({"dram_bytes_per_launch": sum(per) / len(per), "launches_captured": len(per),
                       "grid": grid, "blocks_per_launch": int(grid["value"] // slices) if isinstance(grid, dict) else None,
                       "source_stamp": source_stamp(),
                       "source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/%s_solve_ncu.json" % rnd},
                      open(os.path.join(PROF, rnd + "_solve_traffic.json"), "w"), indent=1)
            print("solve: DRAM bytes per launch", sum(per) / len(per))
        for r in rows:
            print({k: (v["value"] if isinstance(v, dict) else v) for k, v in r.items()})


if __name__ == "__main__":
    main()
