#!/usr/bin/env python
"""Markdown table of the key metrics of every kernel in .ncu-rep files (developer tool).
    python tools/ncu_table.py rep1.ncu-rep [rep2 ...]"""
import csv, io, subprocess, sys
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"),
        ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]
print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|" + "---|" * len(COLS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("CUtensorMap_st, ", "").replace("CUtensorMap_st", "maps")
        name = name.split("(")[0].replace("void ", "")[:60]
        cells = []
        for k, _ in COLS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    f = float(v.replace(",", ""))
                    v = ("%.3f" % f).rstrip("0").rstrip(".") if f < 1000 else "%.0f" % f
                except ValueError:
                    pass
                cells.append(v + (" " + u if u not in ("", "%", "register/thread") else ""))
            else:
                cells.append("-")
        print("| `%s` | " % name + " | ".join(cells) + " |")
