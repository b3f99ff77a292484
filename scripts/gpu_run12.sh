mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -k "refit or gpu_built" > gpurun_out/r02l_pytest_subset.txt 2>&1; tail -4 gpurun_out/r02l_pytest_subset.txt; grep -E "BVH build|rel RMSE|^E  " gpurun_out/r02l_pytest_subset.txt | head -20
B200PT_BVH_BUILDER=gpu python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_lbvh.json 2> gpurun_out/r02l_bench_lbvh.err
python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_lbvh.json')); s=d['roofline']['stages']; print('device-built trees', round(d['value'],1), d['roofline']['model'][-60:], {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
tail -3 gpurun_out/r02l_bench_lbvh.err
