# 8 any-hit candidates per walk (B200PT_KCAND=8: fewer continuation rounds, later bound) against 4, one box
mkdir -p gpurun_out
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02ze_bench_$TAG.json 2> gpurun_out/r02ze_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02ze_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02ze_bench_$TAG.err
}
L=$PWD/vk_gltf_renderer_b200
run kc4 A=1
run kc8 B200PT_LIB=$L/libb200pt_kc8.so
run kc4b A=1
