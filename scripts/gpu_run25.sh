mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_guides.py tests/test_gpu_tonemap.py tests/test_gpu_animation.py -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02za_bench.json 2> gpurun_out/r02za_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02za_bench.json')); s=d['roofline']['stages']; print('default', round(d['value'],1), round(d['e2e']['value'],1), d['config'].get('frame_batch'), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" || tail -3 gpurun_out/r02za_bench.err
