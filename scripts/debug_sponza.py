import sys, os
sys.path.insert(0, '.')
import numpy as np
from vk_gltf_renderer_b200 import synth, hdr
from vk_gltf_renderer_b200.renderer import render_headless, Resources
env = hdr.load_hdr('tests/assets/std_env.hdr')
os.makedirs('gpurun_out', exist_ok=True)
for tag in ('tex', 'notex', 'mip0'):
    scn = synth.synth_sponza(tex_size=256, detail=0.05)
    if tag == 'notex':
        for m in scn.materials:
            m.pbrBaseColorTexture = 0; m.pbrMetallicRoughnessTexture = 0; m.normalTexture = 0
    res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(320, 180))
    kw = dict(ptMaxDepth=6)
    if tag == 'mip0':
        kw['ptTexGradScale'] = 0.0
    pt, img = render_headless(res, 8, **kw)
    np.save('gpurun_out/dbg_%s.npy' % tag, img)
    print(tag, img[..., :3].mean())
