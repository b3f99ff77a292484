# k_shade CTAs of 384 / 448 threads (155 / 144 registers, no spills, 12 / 14 warps per SM) against 512 (128 registers, 16 warps, 44-byte spills)
mkdir -p gpurun_out
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zd_bench_$TAG.json 2> gpurun_out/r02zd_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02zd_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02zd_bench_$TAG.err
}
L=$PWD/vk_gltf_renderer_b200
run s512 A=1
run s384 B200PT_LIB=$L/libb200pt_shade384.so
run s448 B200PT_LIB=$L/libb200pt_shade448.so
run s512b A=1
