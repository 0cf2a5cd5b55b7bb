"""Turn an ncu metrics CSV of tools/profile_run.py (one render, RTB_PIPELINES=1) into the per-ray figures bench.py's
roofline reads from profiles/roofline_traffic.json.

    python tools/ncu_summary.py <metrics.csv> <plain.log> <workload key> [<source note>]   -> prints the JSON entry

metrics.csv : ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,
              smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,
              sm__warps_active.avg.pct_of_peak_sustained_active -k regex:'k_trace|k_shade'
plain.log   : output of the same profile_run.py command without ncu (its `rep 0:` line carries the ray counts)
"""
import csv
import json
import re
import sys


def main():
    path, plain, key = sys.argv[1], sys.argv[2], sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    m = re.search(r"rep 0: .*?\((\d+) extend \+ (\d+) shadow rays\)", open(plain).read())
    rays = int(m.group(1)) + int(m.group(2))
    per = {}
    for r in rows:
        kname = "k_trace" if "k_trace" in r["Kernel Name"] else ("k_shade" if "k_shade" in r["Kernel Name"] else None)
        if not kname:
            continue
        d = per.setdefault(kname, {}).setdefault(r["ID"], {})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    out = {}
    for kname, launches in per.items():
        L = list(launches.values())
        dur_ns = [x["gpu__time_duration.sum"] * ({"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(x.get("unit:gpu__time_duration.sum", "ns"), 1)) for x in L]
        tot = sum(dur_ns)
        w = lambda name: sum(x[name] * d for x, d in zip(L, dur_ns)) / tot if tot else None  # duration-weighted mean
        dram = sum(x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in L)
        l2 = sum(x["lts__t_bytes.sum"] for x in L)
        out[kname] = {"launches": len(L), "time_ms_under_ncu": tot * 1e-6, "dram_bytes": dram, "l2_bytes": l2,
                      "issue_slot_pct": w("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "lanes_per_inst": w("smsp__thread_inst_executed_per_inst_executed.ratio"),
                      "warps_active_pct": w("sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "dram_gbs_under_ncu": dram / tot if tot else None}
    t = out["k_trace"]
    entry = {key: {"k_trace_dram_bytes_per_ray": t["dram_bytes"] / rays, "k_trace_l2_bytes_per_ray": t["l2_bytes"] / rays,
                   "k_trace_issue_slot_pct": t["issue_slot_pct"], "k_trace_lanes_per_inst": t["lanes_per_inst"],
                   "k_trace_warps_active_pct": t["warps_active_pct"], "k_trace_launches_measured": t["launches"],
                   "k_trace_dram_gbs_under_ncu": t["dram_gbs_under_ncu"], "rays_measured": rays,
                   "k_shade": {k: out["k_shade"][k] for k in ("launches", "issue_slot_pct", "lanes_per_inst", "dram_gbs_under_ncu")} if "k_shade" in out else None,
                   "source": note}}
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
