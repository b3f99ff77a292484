# 2 any-hit candidates per walk (the bound shrinks sooner) with 1 .. 4 continuation rounds against 4 candidates / 1 round, one box
mkdir -p gpurun_out
L=$PWD/vk_gltf_renderer_b200
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zg_bench_$TAG.json 2> gpurun_out/r02zg_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02zg_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02zg_bench_$TAG.err
}
run kc4r1 A=1
run kc2r2 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=2
run kc2r3 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=3
run kc2r4 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=4
run kc4r2 B200PT_CONT_ROUNDS=2
run kc4r1b A=1
B200PT_LIB=$L/libb200pt_kc2.so timeout 900 python -m pytest tests/test_golden.py tests/test_gpu_box.py tests/test_gpu_scenes.py -q -m gpu 2>&1 | tail -4
