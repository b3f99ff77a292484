#!/bin/bash
# inside `gpurun --gpus 8`: N = 1 and N = 8 default bench lines of the metric config on ONE box, the driver's invocation (--steps 20 --warmup 5)
TAG=$1
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_n8.json 2> gpurun_out/${TAG}_n8.err
tail -2 gpurun_out/${TAG}_n8.err
python - <<PY
import json
for n in (1, 8):
    try:
        d=json.load(open("gpurun_out/${TAG}_n%d.json" % n)); print("N=%d" % n, d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "batch", d.get("frame_batch"), "launches", d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
