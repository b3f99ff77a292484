#!/bin/bash
# final verification of a round: whole GPU suite, default bench line, ncu launch list of the bench workload, one `--set full` capture
TAG=${1:-r02final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; echo rc=$? >> gpurun_out/${TAG}_pytest_gpu.txt
grep -E "passed|failed|rc=" gpurun_out/${TAG}_pytest_gpu.txt | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.txt | head -20
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("default", d["value"], d["e2e"]["value"], d["roofline"]["frac"], {k:(round(v["share"],3), round(v["ms_per_launch"],4)) for k,v in d["roofline"]["stages"].items()}, d["with_opacity_micromaps"] and d["with_opacity_micromaps"]["value"], d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("bench failed", e)
PY
tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; head -c 600 gpurun_out/${TAG}_bench_reference.json; echo
# launch list of the bench workload (2 wavefronts of 8 frames), one metric, no clock control
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 16 --warmup 0 --profile-only > gpurun_out/${TAG}_launches.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv
# bounce 1 of one frame: k_trace main + continuation + k_shade, full sections with source counters
B200PT_FRAMES_IN_FLIGHT=1 ncu --set full --import-source on --clock-control none -k regex:'k_trace|k_shade' --launch-skip 3 -c 3 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
