# k_shade software pipeline (queue entry two passes ahead, path state prefetched into L2 one pass ahead) against the plain loop, one box
mkdir -p gpurun_out
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-omm-pass > gpurun_out/r02zc_bench_$TAG.json 2> gpurun_out/r02zc_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02zc_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})" 2>/dev/null || tail -3 gpurun_out/r02zc_bench_$TAG.err
}
L=$PWD/vk_gltf_renderer_b200
run prefetch A=1
run plain B200PT_LIB=$L/libb200pt_noprefetch.so
run prefetch2 A=1
run plain2 B200PT_LIB=$L/libb200pt_noprefetch.so
timeout 600 python -m pytest tests/test_gpu_box.py tests/test_golden.py -q -m gpu 2>&1 | tail -3
