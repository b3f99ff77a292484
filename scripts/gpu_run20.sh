mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omm.py tests/test_gpu_materials.py tests/test_gpu_scenes.py -q -m gpu -k "deep_layers or refit or gpu_built or trace_parity or synth_sponza_small" 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-omm-pass > gpurun_out/r02t_bench_$name.json 2> gpurun_out/r02t_bench_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_$name.json')); s=d['roofline']['stages']; print('$name', round(d['value'],1), {k:round(v['ms_per_launch'],4) for k,v in s.items()})"
}
run l2on B200PT_L2_WINDOW=1
run l2off B200PT_L2_WINDOW=0
run l2on_b B200PT_L2_WINDOW=1
run refill18 B200PT_REFILL=18
run refill26 B200PT_REFILL=26
run postpone1 B200PT_POSTPONE=1
run postpone3 B200PT_POSTPONE=3
