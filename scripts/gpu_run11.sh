scripts/gpu_run.sh r02k
BVH_AXISMAP=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_noaxismap.json 2> gpurun_out/r02k_bench_noaxismap.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_noaxismap.json')); s=d['roofline']['stages']; print('axis maps off', round(d['value'],1), d['roofline']['model'][-60:], {k:round(v['ms_per_launch'],4) for k,v in s.items()})
d=json.load(open('gpurun_out/r02k_bench.json')); s=d['roofline']['stages']; print('axis maps on ', round(d['value'],1), d['roofline']['model'][-60:], {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
