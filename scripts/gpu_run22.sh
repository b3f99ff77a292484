# builder A/B on one box: SAH leaf cost (BVH_CPRIM) and bins (BVH_BINS) of the host builder; same kernels
mkdir -p gpurun_out
run() {
  TAG=$1; shift
  env "$@" python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-omm-pass > gpurun_out/r02x_bench_$TAG.json 2> gpurun_out/r02x_bench_$TAG.err
  python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_$TAG.json')); s=d['roofline']['stages']; print('$TAG', round(d['value'],1), d['roofline']['model'][-60:], {k:round(v['ms_per_launch'],4) for k,v in s.items()})" || tail -3 gpurun_out/r02x_bench_$TAG.err
}
run c040 BVH_CPRIM=0.4
run c050 BVH_CPRIM=0.5
run c060 BVH_CPRIM=0.6
run c070 BVH_CPRIM=0.7
run c060b32 BVH_CPRIM=0.6 BVH_BINS=32
run c040_again BVH_CPRIM=0.4
run c060mb8 BVH_CPRIM=0.6 B200PT_LIB=$PWD/vk_gltf_renderer_b200/libb200pt_mb8.so
