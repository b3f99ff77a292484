mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -k "sorted or batching" > gpurun_out/r02i_pytest_subset.txt 2>&1; tail -3 gpurun_out/r02i_pytest_subset.txt
for C in 2 4 5; do
python bench.py --config $C --steps 12 --warmup 3 > gpurun_out/r02i_bench_config$C.json 2> gpurun_out/r02i_bench_config$C.err
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_config$C.json')); print('config $C', d['value'], 'e2e', d['e2e']['value'], d['ms_per_step'], d['rays_per_sample'], d['roofline']['frac'], {k:round(v['share'],3) for k,v in d['roofline']['stages'].items()}, d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)"
done
B200PT_SORT_SHADE=1 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_sort.json 2> gpurun_out/r02i_bench_sort.err
python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_nosort.json 2> gpurun_out/r02i_bench_nosort.err
python -c "
import json
for n in ('sort','nosort'):
    d=json.load(open('gpurun_out/r02i_bench_%s.json'%n)); print(n, d['value'], {k:(round(v['share'],3), round(v['ms_per_launch'],4)) for k,v in d['roofline']['stages'].items()})"
