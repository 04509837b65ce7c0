#!/usr/bin/env python
"""Top SASS instructions of a source-page CSV by executed count and by stall samples (developer tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iex]), int(r[isamp]), r[isrc]))
    except (ValueError, IndexError):
        pass
tot_ex = sum(d[0] for d in data); tot_s = sum(d[1] for d in data)
print("total executed %d, total samples %d, SASS lines %d" % (tot_ex, tot_s, len(data)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("-- by executed")
for ex, sa, src in sorted(data, reverse=True)[:n]:
    print("%6.2f%% ex  %6.2f%% smp  %s" % (100.0 * ex / tot_ex, 100.0 * sa / max(tot_s, 1), src[:110]))
print("-- by samples")
for ex, sa, src in sorted(data, key=lambda d: -d[1])[:n]:
    print("%6.2f%% ex  %6.2f%% smp  %s" % (100.0 * ex / tot_ex, 100.0 * sa / max(tot_s, 1), src[:110]))
