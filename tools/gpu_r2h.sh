#!/bin/bash
# round 2, session h (2 GPUs): sharded bench with the side-stream halo pipeline, parity, config 5
tag=${1:-r2h}
n=${2:-2}
mkdir -p gpurun_out
echo "== bench --gpus $n"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/${tag}_bench${n}.err | grep "^{" | tee gpurun_out/${tag}_bench${n}.json | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench${n}.json").read())
    print("value", d["value"], "ms/step", d["ms_per_step"], "parity", d.get("parity"))
    print(json.dumps(d.get("secondary"), indent=1)[:2500])
    print("e2e", d.get("e2e"))
except Exception as e:
    print("no bench line", e)
PY
tail -5 gpurun_out/${tag}_bench${n}.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $n --steps 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/${tag}_bench${n}_ref.json | cut -c1-500
