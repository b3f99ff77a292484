mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omm.py -q -m gpu 2>&1 | tail -4
for V in libb200pt.so libb200pt_shade640.so libb200pt_shade768.so; do
  B200PT_LIB=$PWD/vk_gltf_renderer_b200/$V python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-omm-pass > gpurun_out/r02w_bench_${V%.so}.json 2> gpurun_out/r02w_bench_${V%.so}.err
  python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_${V%.so}.json')); s=d['roofline']['stages']; print('$V', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
done
