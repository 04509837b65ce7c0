#!/usr/bin/env python
"""Summarises ncu captures into profiles/<tag>_ncu.md (developer tool).

    python tools/ncu_summary.py <tag> [--rep gpurun_out/<tag>_prof.ncu-rep]
                                      [--launches gpurun_out/<tag>_launches.csv]
"""
import argparse
import collections
import csv
import io
import os
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[iname] == "Kernel Name":
            continue
        v = float(r[ival].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit], 1.0)
        agg.setdefault(r[iname].split("(")[0].replace("void ", ""), []).append(v)
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rep = a.rep or "gpurun_out/%s_prof.ncu-rep" % a.tag
    lst = a.launches or "gpurun_out/%s_launches.csv" % a.tag
    lines = ["# ncu summary `%s`" % a.tag, ""]
    if a.note:
        lines += [a.note, ""]
    if os.path.exists(lst):
        agg = launches(lst)
        tot = sum(sum(v) for v in agg.values())
        lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
                  "cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | avg ms | share of listed time |", "|---|---|---|---|"]
        for k, v in agg.items():
            lines.append("| `%s` | %d | %.4f | %.3f |" % (k, len(v), sum(v) / len(v), sum(v) / tot))
        lines.append("")
    if os.path.exists(rep):
        hdr, units, rows = raw_page(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        lines += ["## `ncu --set full --clock-control none` (one launch per kernel)", ""]
        for r in rows:
            lines += ["### `%s`" % r[idx["Kernel Name"]].split("(")[0].replace("void ", ""), "",
                      "| metric | value |", "|---|---|"]
            for m, label in METRICS:
                if m in idx and r[idx[m]] not in ("", "n/a"):
                    lines.append("| %s (`%s`) | %s %s |" % (label, m, r[idx[m]], units[idx[m]]))
            rd, wr = idx.get("dram__bytes_read.sum"), idx.get("dram__bytes_write.sum")
            if rd is not None and wr is not None:
                def gb(i):
                    v = float(r[i].replace(",", ""))
                    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(units[i], 1.0)
                lines.append("| **traffic** (read + write) | %.3f GB |" % (gb(rd) + gb(wr)))
            lines.append("")
    os.makedirs("profiles", exist_ok=True)
    path = "profiles/%s_ncu.md" % a.tag
    open(path, "w").write("\n".join(lines))
    print(open(path).read())


if __name__ == "__main__":
    main()
