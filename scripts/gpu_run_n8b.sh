#!/bin/bash
# inside `gpurun --gpus 8`: the default bench at N=1 and N=8 on the same box, nothing else (box time is charged 8x)
TAG=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_omm.py -q -m gpu -k cpp_host -s 2>&1 | grep -E "useOpacityMicromap|passed|failed" | head -5
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/${TAG}_n8.json 2> gpurun_out/${TAG}_n8.err
tail -2 gpurun_out/${TAG}_n8.err
python - <<PY
import json
for n in (1, 8):
    try:
        d=json.load(open("gpurun_out/${TAG}_n%d.json" % n))
        print("N=%d" % n, d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "batch", d.get("frame_batch"), "omm", d["with_opacity_micromaps"] and d["with_opacity_micromaps"]["value"], d["clocks"])
    except Exception as e:
        print("N=%d failed" % n, e)
PY
