#!/bin/bash
# inside `gpurun --gpus 8`: the N = 8 bench lines of config 3 (1080p, the metric config) and config 5 (4K), and a per-rank timeline
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
tail -2 gpurun_out/${TAG}_n${N}.err
$TR --master-port 29513 bench.py --gpus $N --config 5 --steps 20 --warmup 3 > gpurun_out/${TAG}_n${N}_config5.json 2> gpurun_out/${TAG}_n${N}_config5.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_n${N}.json", "gpurun_out/${TAG}_n${N}_config5.json"):
    try:
        d=json.load(open(f)); print(f, d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "batch", d.get("frame_batch"), "lanes", d.get("frames_in_flight"), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
B200PT_TIMELINE=$PWD/gpurun_out/${TAG}_timeline_n${N}.csv $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 20 --profile-only > gpurun_out/${TAG}_timeline_n${N}.log 2>&1
for f in gpurun_out/${TAG}_timeline_n${N}.csv.*; do python scripts/timeline_summary.py $f 1 | head -4; done
