#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under
profiles/ (the .ncu-rep itself is scratch).  Runs in the build container (ncu -i
needs no GPU).

    python tools/ncu_summary.py --tag r01 --rep gpurun_out/prof.ncu-rep --launches gpurun_out/launches.csv
"""
import argparse
import csv
import io
import json
import os
import subprocess
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
       "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def summarise_rep(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in RAW:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": to_float(r[i]), "unit": units[i]}
        res.append(d)
    return res


def bytes_of(m):
    v, u = m["value"], m["unit"].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult


def summarise_launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = to_float(r[vi])
        if isinstance(v, float):
            name = r[ki]
            short = name if len(name) < 100 else name[:97] + "..."
            agg[short][0] += 1
            agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    return [{"kernel": k, "launches": n, "total_us": t / 1e3, "avg_us": t / 1e3 / n, "share": t / tot}
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]


def summarise_dram(path, pattern="bnnp_step_kernel"):
    """Per-launch DRAM bytes of consecutive launches profiled in ONE pass with
    `--cache-control none` (no replay, no flush: the L2 state is the one the previous
    launch left, as in the timed region of bench.py)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = defaultdict(dict)
    for r in rows[1:]:
        if pattern in r[ki]:
            v = to_float(r[vi])
            if isinstance(v, float):
                mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0,
                        "nsecond": 1e-3, "msecond": 1e3}.get(r[ui].lower(), 1)
                per[r[ii]][r[mi]] = v * mult
    ls = [d for d in per.values() if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d]
    n = len(ls)
    return {"launches": n,
            "dram_bytes_read_per_launch": sum(d["dram__bytes_read.sum"] for d in ls) / n,
            "dram_bytes_write_per_launch": sum(d["dram__bytes_write.sum"] for d in ls) / n,
            "dram_bytes_per_launch": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ls) / n,
            "us_per_launch_under_ncu": sum(d.get("gpu__time_duration.sum", 0.0) for d in ls) / n,
            "per_launch_mb": [round((d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / 1e6, 1) for d in ls]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--dram-csv", help="single-pass dram__bytes capture with --cache-control none")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.rep:
        ks = summarise_rep(a.rep)
        mine = [k for k in ks if "bnnp_step_kernel" in k["kernel"]]
        summary = {"note": a.note, "source": os.path.basename(a.rep), "kernels": ks}
        if mine:
            tr = [bytes_of(k["dram__bytes_read.sum"]) + bytes_of(k["dram__bytes_write.sum"]) for k in mine]
            summary["dram_bytes_per_launch"] = sum(tr) / len(tr)
            summary["dram_bytes_read_per_launch"] = sum(bytes_of(k["dram__bytes_read.sum"]) for k in mine) / len(mine)
            summary["dram_bytes_write_per_launch"] = sum(bytes_of(k["dram__bytes_write.sum"]) for k in mine) / len(mine)
        with open(os.path.join(ROOT, "profiles", f"{a.tag}_ncu_step_kernel.json"), "w") as f:
            json.dump(summary, f, indent=1)
        # bench.py reads the latest capture's traffic from this fixed name
        with open(os.path.join(ROOT, "profiles", "ncu_step_kernel.json"), "w") as f:
            json.dump({k: v for k, v in summary.items() if k != "kernels"} | {"tag": a.tag}, f, indent=1)
        for k in mine[:1]:
            print(json.dumps({m: v["value"] for m, v in k.items() if isinstance(v, dict)}, indent=1))
    if a.dram_csv:
        d = summarise_dram(a.dram_csv)
        d["note"] = ("steady state: consecutive launches, one ncu pass, --cache-control none (L2 as the previous launch "
                     "left it). " + a.note)
        with open(os.path.join(ROOT, "profiles", f"{a.tag}_ncu_dram_steady.json"), "w") as f:
            json.dump(d, f, indent=1)
        fixed = os.path.join(ROOT, "profiles", "ncu_step_kernel.json")
        cur = json.load(open(fixed)) if os.path.exists(fixed) else {}
        if "dram_bytes_per_launch" in cur and "dram_bytes_per_launch_cold" not in cur:
            cur["dram_bytes_per_launch_cold"] = cur["dram_bytes_per_launch"]      # the --set full capture flushes caches
        cur.update({"dram_bytes_per_launch": d["dram_bytes_per_launch"], "steady_tag": a.tag,
                    "dram_bytes_read_per_launch": d["dram_bytes_read_per_launch"],
                    "dram_bytes_write_per_launch": d["dram_bytes_write_per_launch"]})
        with open(fixed, "w") as f:
            json.dump(cur, f, indent=1)
        print(json.dumps({k: v for k, v in d.items() if k != "note"}))
    if a.launches:
        ls = summarise_launches(a.launches)
        with open(os.path.join(ROOT, "profiles", f"{a.tag}_ncu_launches.json"), "w") as f:
            json.dump({"note": a.note, "source": os.path.basename(a.launches), "kernels": ls}, f, indent=1)
        for k in ls[:6]:
            print(f'{k["share"]*100:5.1f}%  {k["launches"]:4d} x {k["avg_us"]:8.1f} us  {k["kernel"][:80]}')


if __name__ == "__main__":
    main()
