#!/bin/bash
# usage: scripts/gpu_run_n.sh TAG N   (inside `gpurun --gpus N`): N=1 and N-rank bench lines on the same box + a short N-rank timeline
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
        print("N=%d" % n, d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "batch", d.get("frame_batch"), "lanes", d.get("frames_in_flight"), "launches", d["gpu_launches"])
    except Exception as e:
        print("N=%d failed" % n, e)
PY
B200PT_TIMELINE=$PWD/gpurun_out/${TAG}_timeline_n${N}.csv python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 32 --warmup 16 --profile-only > gpurun_out/${TAG}_timeline_n${N}.log 2>&1
for f in gpurun_out/${TAG}_timeline_n${N}.csv.*; do python scripts/timeline_summary.py $f 1 | head -12; done
