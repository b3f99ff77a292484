import os, sys, time
sys.path.insert(0, '.')
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for p in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us', '/sys/fs/cgroup/cpu/cpu.cfs_period_us'):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, 'n/a')
import numpy as np
from vk_gltf_renderer_b200 import synth, hdr, camera as cm
from oracle import oracle as O
env = hdr.load_hdr('tests/assets/std_env.hdr')
scn = synth.synth_sponza(tex_size=256, detail=0.3)
o = O.Oracle(); o.set_scene(scn); o.set_environment(env)
W, H = 640, 360
fi = cm.make_frame_info(scn.camera, W, H)
for th in (1, 8, 32, 64, 128):
    acc = np.zeros((H, W, 4), np.float32)
    pc = cm.make_push_constant(scn.camera, H, frame_count=0, total_samples=0, max_depth=12)
    t = time.time(); o.render_frame(fi, pc, acc, threads=th); dt = time.time() - t
    print('threads', th, round(dt, 2), 's')
