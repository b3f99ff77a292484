mkdir -p gpurun_out
for G in 1 0; do
B200PT_CONT_SMALL_GRID=$G python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zj_bench_small$G.json 2> gpurun_out/r02zj_bench_small$G.err
python -c "
import json; d=json.load(open('gpurun_out/r02zj_bench_small$G.json')); s=d['roofline']['stages']; print('small$G', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['share'],3) for k,v in s.items()})" || tail -3 gpurun_out/r02zj_bench_small$G.err
done
