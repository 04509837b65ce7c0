#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
echo "== conv tests"; timeout 600 python -m pytest tests/test_conv3x3.py -m gpu -q -x 2>&1 | tail -2
echo "== convs bench"; timeout 600 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl | cut -c1-200
echo "== launch list of the forwards"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 2 --warmup 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2g_launches.csv")) if len(r) > 8]
hdr = rows[0]
i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[1:]:
    try:
        seq.append((r[i_name], float(r[i_val].replace(",", ""))))
    except ValueError:
        pass
# the last forward = everything after the third-last splat launch group: take the last quarter by splat count
idx = [i for i, (n, _) in enumerate(seq) if "splat_fwd" in n]
start = idx[-4] if len(idx) >= 8 else 0
# walk back to the layout-change kernel that opens the forward
for j in range(start, -1, -1):
    if "nchw_to_nhwc" in seq[j][0]:
        start = j
        break
last = seq[start:]
agg = collections.OrderedDict()
for n, v in last:
    a = agg.setdefault(n[:86], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("last forward: %d launches, %.2f ms of kernel time (ncu, serialised)" % (len(last), tot / 1e6))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% %4d  %s" % (v / 1e3, 100 * v / tot, n, k))
PY
