# frame batch / frames in flight at the DRIVER's own invocation (--steps 20 --warmup 5), leaf cost 0.7, 7 vs 8 walk CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_animation.py tests/test_gpu_tonemap.py -q -m gpu 2>&1 | tail -15
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02y_bench_$TAG.json 2> gpurun_out/r02y_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02y_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" || tail -3 gpurun_out/r02y_bench_$TAG.err
}
MB8=$PWD/vk_gltf_renderer_b200/libb200pt_mb8.so
run b8l4 BVH_CPRIM=0.7
run b10l4 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=10
run b7l4 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=7
run b5l4 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=5
run b4l5 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=4 B200PT_FRAMES_IN_FLIGHT=5
run b5l4mb8 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=5 B200PT_LIB=$MB8
run b8l4mb8 BVH_CPRIM=0.7 B200PT_LIB=$MB8
run b7l4mb8 BVH_CPRIM=0.7 B200PT_FRAME_BATCH=7 B200PT_LIB=$MB8
run b10l4mb8c10 BVH_CPRIM=1.0 B200PT_FRAME_BATCH=10 B200PT_LIB=$MB8
run b8l4mb8c10 BVH_CPRIM=1.0 B200PT_LIB=$MB8
