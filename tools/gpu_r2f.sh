#!/bin/bash
# round 2, session f: quad-transpose stores, glue kernels; launch list of one config-3 forward
tag=${1:-r2f}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.txt
echo "== chains v3 bench"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3.jsonl | cut -c1-330
echo "== convs bench"; timeout 600 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl | cut -c1-330
echo "== cfg3 forward"
timeout 600 python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 5 --warmup 2 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3.json | cut -c1-900
echo "== launch list of one forward"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 2 --warmup 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2f_launches.csv")) if len(r) > 8]
hdr = rows[0]
i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[i_val].replace(",", ""))
    except ValueError:
        continue
    k = r[i_name][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% %4d  %s" % (v / 1e3, 100 * v / tot, n, k))
PY
