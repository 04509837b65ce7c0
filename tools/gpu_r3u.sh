#!/bin/bash
# final state of round 2: smoke, full GPU suite, bench.py (N=1), tile reader, training end to end
tag=${TAG:-r3u}
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke ok"
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/${tag}_pytest.txt
echo "== bench"; timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" > gpurun_out/${tag}_bench.json
python - <<'PY'
import json, os
tag = os.environ.get("TAG", "r3u")
d = json.loads(open("gpurun_out/%s_bench.json" % tag).read().strip().splitlines()[-1])
sec = d.pop("secondary", {})
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"]["reasons"])
print({k: (round(v["ms"], 2) if isinstance(v, dict) and "ms" in v else None) for k, v in sec.get("config4_train_step", {}).items()})
print(sec.get("config3_forward", {}).get("ms"), sec.get("error_config34"))
PY
echo "== tiles"; timeout 300 python benchmarks/tiles_bench.py 2>&1 | grep "^{" | tee gpurun_out/${tag}_tiles_bench.json | cut -c300-520
echo "== train e2e"
for d in 16 32; do timeout 500 python benchmarks/train_e2e_bench.py --prefetch $d --steps 64 2>&1 | grep -E "^\{" | tee -a gpurun_out/${tag}_train_e2e.jsonl | cut -c300-1100; done
