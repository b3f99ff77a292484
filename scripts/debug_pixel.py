"""Print one pixel's path from the CUDA debug build (run on GPU) or the oracle (--oracle)."""
import sys, os, ctypes as C
sys.path.insert(0, '.')
import numpy as np
from vk_gltf_renderer_b200 import synth, hdr, camera as cm, _lib
px = [tuple(map(int, a.split(','))) for a in sys.argv[1].split(';')]
depth = int(sys.argv[2])
use_oracle = len(sys.argv) > 3 and sys.argv[3] == '--oracle'
env = hdr.load_hdr('tests/assets/std_env.hdr')
if os.environ.get('DBG_SCENE') == 'glass':
    scn = synth.synth_glass(n=48, scatter=True)
    W, H = 128, 128
else:
    scn = synth.synth_sponza(tex_size=64, detail=0.05)
    for m in scn.materials:
        m.pbrBaseColorTexture = 0; m.pbrMetallicRoughnessTexture = 0; m.normalTexture = 0
    W, H = 640, 360
fi = cm.make_frame_info(scn.camera, W, H)
if use_oracle:
    from oracle import oracle as O
    o = O.Oracle(); o.set_scene(scn); o.set_environment(env)
    for (x, y) in px:
        pc = cm.make_push_constant(scn.camera, H, frame_count=0, total_samples=0, max_depth=depth)
        pc.mouseCoord[:] = [x, y]
        acc = np.zeros((1, W, 4), np.float32)
        print('=== pixel', x, y, file=sys.stderr); sys.stderr.flush()
        o.render_frame(fi, pc, acc, y0=y, rows=1, threads=1)
        print('result', acc[0, x], file=sys.stderr)
else:
    from vk_gltf_renderer_b200.renderer import PathTracer, Resources
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libb200pt_debug.so')
    res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(W, H))
    pt = PathTracer(0); pt.onAttach(res)
    for (x, y) in px:
        pc = cm.make_push_constant(scn.camera, H, frame_count=0, total_samples=0, max_depth=depth)
        pc.mouseCoord[:] = [x, y]
        print('=== pixel', x, y, flush=True)
        pt.render_frame_raw(fi, pc); pt.synchronize()
        print('result', pt.read_accum()[y, x], flush=True)
