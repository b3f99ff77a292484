scripts/gpu_run.sh r02g
for B in 2 4; do for L in 2 4; do
B200PT_FRAME_BATCH=$B B200PT_FRAMES_IN_FLIGHT=$L python bench.py --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/r02g_bench_b${B}_l${L}.json 2> gpurun_out/r02g_bench_b${B}_l${L}.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_b${B}_l${L}.json')); print('batch $B lanes $L', d['value'], d['e2e']['value'], d['gpu_launches'])"
done; done
