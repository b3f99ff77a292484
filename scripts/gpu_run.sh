#!/bin/bash
# one gpurun call: full GPU test suite, the default bench line, quick bench lines of the tuning variants, a light ncu pass.
# usage: scripts/gpu_run.sh TAG [variant.so ...]
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; echo rc=$? >> gpurun_out/${TAG}_pytest_gpu.txt
grep -E "passed|failed|rc=" gpurun_out/${TAG}_pytest_gpu.txt | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.txt | head -20
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("default", d["value"], d["e2e"]["value"], d["roofline"]["frac"], {k:(round(v["share"],3), round(v["ms_per_launch"],4)) for k,v in d["roofline"]["stages"].items()})
except Exception as e:
    print("bench failed", e)
PY
for V in "$@"; do
  B200PT_LIB=$PWD/vk_gltf_renderer_b200/$V python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${V%.so}.json 2> gpurun_out/${TAG}_bench_${V%.so}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${V%.so}.json"))
    print("$V", d["value"], {k:(round(v["share"],3), round(v["ms_per_launch"],4)) for k,v in d["roofline"]["stages"].items()})
except Exception as e:
    print("$V failed", e)
PY
done
if [ -n "$NCU_LIGHT" ]; then
  B200PT_FRAMES_IN_FLIGHT=1 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_(trace|shadow|shade|alpha)' -c 60 --csv --log-file gpurun_out/${TAG}_ncu_light.csv python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/${TAG}_ncu_light.log 2>&1
fi
if [ -n "$TIMELINE" ]; then
  B200PT_TIMELINE=$PWD/gpurun_out/${TAG}_timeline.csv python bench.py --steps 12 --warmup 4 --profile-only > gpurun_out/${TAG}_timeline.log 2>&1
  for f in gpurun_out/${TAG}_timeline.csv.*; do python scripts/timeline_summary.py $f 6; done
fi
if [ -n "$NCU_FULL" ]; then
  # one bounce-1 launch of the kernels named in $NCU_FULL (regex), full sections + source counters
  B200PT_FRAMES_IN_FLIGHT=1 ncu --set full --import-source on --clock-control none -k regex:"$NCU_FULL" --launch-skip ${NCU_SKIP:-2} -c ${NCU_COUNT:-1} -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/${TAG}_full.ncu-rep
fi
