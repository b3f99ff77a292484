mkdir -p gpurun_out
for V in libb200pt.so libb200pt_mb8.so libb200pt_shade256.so libb200pt_mb6.so; do
B200PT_LIB=$PWD/vk_gltf_renderer_b200/$V python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/r02p_bench_${V%.so}.json 2> gpurun_out/r02p_bench_${V%.so}.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_${V%.so}.json')); s=d['roofline']['stages']; print('$V', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
done
