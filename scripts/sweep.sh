#!/bin/bash
for mb in 4 5 6; do
  B200PT_LIB=$PWD/vk_gltf_renderer_b200/libb200pt_mb$mb.so python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); st=d['roofline']['stages']
print('shade minBlocks $mb  value %.1f  trace %.3f shade %.3f post %.3f'%(d['value'],st['k_trace']['ms_per_launch'],st['k_shade']['ms_per_launch'],st['k_post']['ms_per_launch']))"
done
