mkdir -p gpurun_out
for BL in "4 4" "8 2" "8 4" "6 4" "4 6" "3 6"; do set -- $BL
B200PT_FRAME_BATCH=$1 B200PT_FRAMES_IN_FLIGHT=$2 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02m_bench_b$1_l$2.json 2> gpurun_out/r02m_bench_b$1_l$2.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_b$1_l$2.json')); print('batch $1 lanes $2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'per-frame-read', round(d['e2e']['per_frame_readback_value'],1))"
done
