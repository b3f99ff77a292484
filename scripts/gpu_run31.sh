mkdir -p gpurun_out
L=$PWD/vk_gltf_renderer_b200
B200PT_LIB=$L/libb200pt_kc2.so timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_box.py -q -m gpu 2>&1 | tail -4
B200PT_LIB=$L/libb200pt_kc2.so timeout 900 python -m pytest tests/test_gpu_scenes.py -q -m gpu -k "soup or atrium or sponza or deep or layers" 2>&1 | tail -6
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zf_bench_$TAG.json 2> gpurun_out/r02zf_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02zf_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02zf_bench_$TAG.err
}
run kc2 B200PT_LIB=$L/libb200pt_kc2.so
run kc4 A=1
run kc2b B200PT_LIB=$L/libb200pt_kc2.so
