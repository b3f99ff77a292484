import sys, os
sys.path.insert(0, '.')
import numpy as np
from vk_gltf_renderer_b200 import synth, hdr
from vk_gltf_renderer_b200.renderer import render_headless, Resources
env = hdr.load_hdr('tests/assets/std_env.hdr')
scn = synth.synth_sponza(tex_size=64, detail=0.05)
for m in scn.materials:
    m.pbrBaseColorTexture = 0; m.pbrMetallicRoughnessTexture = 0; m.normalTexture = 0
for depth in (1, 2, 3, 6):
    res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(640, 360))
    pt, img = render_headless(res, 1, ptMaxDepth=depth)
    np.save('gpurun_out/dbg_depth%d.npy' % depth, img)
    print(depth, pt.stats())
