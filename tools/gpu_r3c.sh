#!/bin/bash
# round 2, training pipeline: tests, kernel times, config-4 variants, ncu of the new kernels, launch list of one step
tag=${TAG:-r3c}
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_train_ops.py tests/test_train_pipeline.py tests/test_conv3x3.py tests/test_optim.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.txt
echo "== kernels"; timeout 600 python benchmarks/train_kernels_bench.py 2>&1 | grep "^{" | cut -c1-230
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_train_ops.py tests/test_train_pipeline.py -m gpu -q -x -k "not cuda_graph" 2>&1 | tail -4 | tee gpurun_out/${tag}_sanitizer.txt
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wgrad|linear_kernel|conv3x3_kernel|nchw_to_nhwc' -c 36 -f \
  -o gpurun_out/${tag}_train_prof python benchmarks/train_kernels_bench.py --ncu > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_ncu.log
echo "== launch list of one eager step"
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
tag = os.environ.get("TAG", "r3c")
rows = [r for r in csv.reader(open("gpurun_out/%s_launches.csv" % tag)) if len(r) > 10 and r[0].isdigit()]
half = rows                                       # cudaProfilerStart/Stop bracket one warm step
agg = collections.OrderedDict()
for r in half:
    name = r[4].split("(")[0].replace("void ", "")[:70]
    t = float(r[-1].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
repo = sum(a[1] for k, a in agg.items()
           if k.startswith(("sbmc::", "lin::", "c3::", "wg::", "tr::", "wb::")))
with open("gpurun_out/%s_step_launch_list.txt" % tag, "w") as f:
    f.write("one config-4 training step (bf16 pipeline, eager), ncu gpu__time_duration per kernel (serialised, cold)\n")
    f.write("%d launches, %.2f ms of kernel time; repo kernels: %.1f %% of it\n" % (len(half), tot / 1e6, 100 * repo / tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%5d x %-72s %8.3f ms %5.1f %%\n" % (a[0], k, a[1] / 1e6, 100 * a[1] / tot))
print(open("gpurun_out/%s_step_launch_list.txt" % tag).read()[:3000])
PY
echo "== config 4"
timeout 900 python - <<'PY' 2>&1 | grep -v Warning | tail -6
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch as th
import bench
class A: pass
sec = bench.run_secondary(A(), th, None, th.device("cuda", 0), 0, 1)
c4 = sec["config4_train_step"]
print({k: round(v["ms"], 2) for k, v in c4.items() if isinstance(v, dict)})
print(sec.get("error_config34"))
open("gpurun_out/%s_secondary.json" % os.environ.get("TAG", "r3c"), "w").write(json.dumps(sec))
PY
