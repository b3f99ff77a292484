mkdir -p gpurun_out
for WS in "8 2" "7 2" "8 1" "7 1" "14 2" "7 3"; do set -- $WS
B200PT_WALK_GRID=$1 B200PT_SHADE_GRID=$2 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02q_bench_w$1_s$2.json 2> gpurun_out/r02q_bench_w$1_s$2.err
python -c "
import json; d=json.load(open('gpurun_out/r02q_bench_w$1_s$2.json')); s=d['roofline']['stages']; print('walk grid $1/SM shade grid $2/SM', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
done
