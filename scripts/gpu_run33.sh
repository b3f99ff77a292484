# continuation rounds sweep: 4 candidates x {2,3,4,6} rounds, 2 candidates x {4,5,6,8}
mkdir -p gpurun_out
L=$PWD/vk_gltf_renderer_b200
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zh_bench_$TAG.json 2> gpurun_out/r02zh_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02zh_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['share'],3) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02zh_bench_$TAG.err
}
run kc4r2 B200PT_CONT_ROUNDS=2
run kc4r3 B200PT_CONT_ROUNDS=3
run kc4r4 B200PT_CONT_ROUNDS=4
run kc4r6 B200PT_CONT_ROUNDS=6
run kc2r4 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=4
run kc2r5 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=5
run kc2r6 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=6
run kc2r8 B200PT_LIB=$L/libb200pt_kc2.so B200PT_CONT_ROUNDS=8
