#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` capture of one batch step: per kernel launch the DRAM bytes
(read + write), the warp instructions executed, issue-slot and pipe utilisation and the duration alone under ncu.
bench.py reads it for `roofline.traffic` and for the issue-ceiling figures (instructions per launch are a property of
the workload, not of the run).  usage: ncu_traffic.py out.json report.ncu-rep"""
import csv
import io
import json
import re
import subprocess
import sys

MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "inst": 1, "%": 1}


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", n)
    return m.group(1) + (m.group(2) or "")


def main(out, rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * MUL.get(units[col[k]], 1)

    launches, seen = [], {}
    for r in data:
        base = short(r[col["Kernel Name"]])
        if base == "k_unpack" and seen.get("k_unpack"):
            break                                   # the capture ran into the next step: one step only
        seen[base] = seen.get(base, 0) + 1
        launches.append((base, seen[base], r))
    res = {"source": rep.split("/")[-1], "launches": [], "kernels": {}}
    for base, k, r in launches:
        key = base if seen[base] == 1 else "%s#%d" % (base, k)
        e = {"kernel": key, "grid": r[col["Grid Size"]].replace(" ", ""),
             "ncu_duration_us": val(r, "gpu__time_duration.sum"),
             "dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
             "warp_inst_per_launch": val(r, "smsp__inst_executed.sum"),
             "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "pipe_alu_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             "pipe_fma_pct": val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             "pipe_xu_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
             "pipe_lsu_pct": val(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
             "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
        res["launches"].append(e)
        res["kernels"][key] = e
    res["step_warp_inst"] = sum(e["warp_inst_per_launch"] for e in res["launches"])
    res["step_ncu_duration_us"] = sum(e["ncu_duration_us"] for e in res["launches"])
    res["step_dram_bytes"] = sum(e["dram_bytes_per_launch"] for e in res["launches"])
    json.dump(res, open(out, "w"), indent=1)
    for e in res["launches"]:
        print("%-22s %-14s %8.1f us %7.1f M inst  issue %5.1f %%  dram %7.1f MB" % (e["kernel"], e["grid"], e["ncu_duration_us"],
              e["warp_inst_per_launch"] / 1e6, e["issue_active_pct"], e["dram_bytes_per_launch"] / 1e6))
    print("step: %.1f us alone, %.1f M warp instructions, %.1f MB DRAM" % (res["step_ncu_duration_us"], res["step_warp_inst"] / 1e6,
                                                                          res["step_dram_bytes"] / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
