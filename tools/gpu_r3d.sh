#!/bin/bash
# final-state checks: GPU tests of the scripts, bench.py (N=1) as the driver runs it, reference arm, launch list of one training step
tag=${TAG:-r3d}
mkdir -p gpurun_out
echo "== script + capi tests"; timeout 900 python -m pytest tests/test_scripts.py tests/test_capi.py tests/test_train_pipeline.py -m gpu -q 2>&1 | tail -3
echo "== bench.py"; timeout 1200 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" > gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
python - <<'PY'
import json, os
tag = os.environ.get("TAG", "r3d")
d = json.loads(open("gpurun_out/%s_bench.json" % tag).read().strip().splitlines()[-1])
sec = d.pop("secondary", {})
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
print({k: (round(v["ms"], 2) if isinstance(v, dict) and "ms" in v else None) for k, v in sec.get("config4_train_step", {}).items()})
print(sec.get("config3_forward", {}).get("ms"), sec.get("error_config34"))
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" | cut -c1-400 | tee gpurun_out/${tag}_bench_ref.json
echo "== launch list of one eager training step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv \
  python - > gpurun_out/${tag}_launch.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch as th
from sbmc_b200 import interfaces, models
dev = th.device("cuda", 0)
th.manual_seed(0)
net = models.Multisteps(93, 3).to(dev).train()
net.bf16_train = True
iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True)
batch = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev), "features": th.randn(8, 8, 93, 128, 128, device=dev),
         "global_features": th.randn(8, 3, 1, 1, device=dev), "target_image": th.rand(8, 3, 128, 128, device=dev)}
for _ in range(2):
    iface.train_step(batch)
th.cuda.synchronize()
th.cuda.profiler.start()
iface.train_step(batch)
th.cuda.synchronize()
th.cuda.profiler.stop()
PY
python - <<'PY'
import csv, collections, os
tag = os.environ.get("TAG", "r3d")
rows = [r for r in csv.reader(open("gpurun_out/%s_launches.csv" % tag)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")[:70]
    t = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
mine = ("sbmc::", "lin::", "c3::", "wg::", "tr::", "wb::")
repo = sum(a[1] for k, a in agg.items() if k.startswith(mine))
nrepo = sum(a[0] for k, a in agg.items() if k.startswith(mine))
with open("gpurun_out/%s_step_launch_list.txt" % tag, "w") as f:
    f.write("one config-4 training step (bf16 pipeline, eager; cudaProfilerStart/Stop around the third step), ncu gpu__time_duration per kernel (serialised, cold)\n")
    f.write("%d launches (%d of them repo kernels), %.2f ms of kernel time; repo kernels: %.1f %% of it\n" % (len(rows), nrepo, tot / 1e6, 100 * repo / tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%5d x %-72s %8.3f ms %5.1f %%\n" % (a[0], k, a[1] / 1e6, 100 * a[1] / tot))
print(open("gpurun_out/%s_step_launch_list.txt" % tag).read()[:2500])
PY
