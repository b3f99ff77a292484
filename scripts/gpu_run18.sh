mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omm.py tests/test_gpu_scenes.py -q -m gpu -k "omm or micromap or trace_parity or deep_anyhit or direction_bucketed" -rA 2>&1 | tail -25 > gpurun_out/r02r_pytest_omm.txt
cat gpurun_out/r02r_pytest_omm.txt | tail -22
for O in 0 4 5; do
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --omm $O > gpurun_out/r02r_bench_omm$O.json 2> gpurun_out/r02r_bench_omm$O.err
python -c "
import json; d=json.load(open('gpurun_out/r02r_bench_omm$O.json')); s=d['roofline']['stages']; print('omm level $O', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()}, {k:round(v['share'],3) for k,v in s.items()})"
done
B200PT_SORT_RAYS=1 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02r_bench_sortrays.json 2> gpurun_out/r02r_bench_sortrays.err
python -c "
import json; d=json.load(open('gpurun_out/r02r_bench_sortrays.json')); s=d['roofline']['stages']; print('sort rays', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
