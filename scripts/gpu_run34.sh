mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zi_bench.json 2> gpurun_out/r02zi_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02zi_bench.json')); s=d['roofline']['stages']; print('smallgrid', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['share'],3) for k,v in s.items()})" || tail -3 gpurun_out/r02zi_bench.err
timeout 100 python -m pytest tests/test_golden.py tests/test_gpu_box.py -q -m gpu 2>&1 | tail -2
timeout 100 python -m pytest tests/test_gpu_scenes.py -q -m gpu -k "layers or deep or stacked" 2>&1 | tail -2
