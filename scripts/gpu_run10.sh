mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -k "plane" > gpurun_out/r02j_pytest_subset.txt 2>&1; tail -3 gpurun_out/r02j_pytest_subset.txt; grep -E "rel RMSE|^E " gpurun_out/r02j_pytest_subset.txt | head
for R in 18 22 26 30; do for P in 1 2 3; do
B200PT_REFILL=$R B200PT_POSTPONE=$P python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_r${R}_p${P}.json 2> gpurun_out/r02j_bench_r${R}_p${P}.err
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_r${R}_p${P}.json')); s=d['roofline']['stages']; print('refill $R postpone $P', round(d['value'],1), round(s['k_trace']['ms_per_launch'],4), round(s['k_shadow']['ms_per_launch'],4))"
done; done
