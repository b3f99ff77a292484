mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omm.py -q -m gpu -rA 2>&1 | tail -60 > gpurun_out/r02s_pytest_omm.txt
grep -E "PASSED|FAILED|ERROR|passed|failed|Error|assert" gpurun_out/r02s_pytest_omm.txt | head -40
python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02s_bench_default.json 2> gpurun_out/r02s_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_default.json')); s=d['roofline']['stages']; print('default', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()}, d['with_opacity_micromaps'])"
tail -3 gpurun_out/r02s_bench_default.err
python bench.py --steps 24 --warmup 4 --no-cpu-baseline --omm 6 > gpurun_out/r02s_bench_omm6.json 2> gpurun_out/r02s_bench_omm6.err
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_omm6.json')); s=d['roofline']['stages']; print('omm 6', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()}, d['roofline']['frac'])"
