# new GPU tests (animation feed incl. the device-side node hierarchy, tonemap), then A/B on one box: the default build (leaf cost 0.7,
# 8 walk CTAs / SM, compile-time walk protocol) against 7 CTAs / SM and against the runtime-flag walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_animation.py tests/test_gpu_tonemap.py -q -m gpu 2>&1 | tail -15
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02z_bench_$TAG.json 2> gpurun_out/r02z_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02z_bench_$TAG.err
}
L=$PWD/vk_gltf_renderer_b200
run default A=1
run mb7 B200PT_LIB=$L/libb200pt_mb7.so
run rtmode B200PT_LIB=$L/libb200pt_rtmode.so
run default2 A=1
run c10 BVH_CPRIM=1.0
run b10 B200PT_FRAME_BATCH=10
