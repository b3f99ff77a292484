#!/bin/bash
# inside `gpurun --gpus N`: N=1 and N-rank default bench lines on the same box (plus the reference arm at N ranks: rank 0 only works)
TAG=$1; N=$2
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
tail -3 gpurun_out/${TAG}_n${N}.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.load(open("gpurun_out/${TAG}_n%d.json" % n))
        print("N=%d" % n, d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "batch", d.get("frame_batch"), "omm", d["with_opacity_micromaps"] and d["with_opacity_micromaps"]["value"])
    except Exception as e:
        print("N=%d failed" % n, e)
PY
