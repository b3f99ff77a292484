import sys, os
sys.path.insert(0, '.')
import numpy as np
from vk_gltf_renderer_b200 import synth, hdr
from vk_gltf_renderer_b200.renderer import render_headless, Resources
env = hdr.load_hdr('tests/assets/std_env.hdr')
scn = synth.synth_glass(n=48, scatter=True)
for frames in (1, 8):
    res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(128, 128))
    pt, img = render_headless(res, frames, ptMaxDepth=12)
    np.save('gpurun_out/dbg_glass%d.npy' % frames, img)
    print(frames, pt.stats())
