import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from vk_gltf_renderer_b200 import synth, hdr
from vk_gltf_renderer_b200.renderer import PathTracer, Resources
from oracle import oracle as O
from gpu_util import random_rays
env = hdr.load_hdr('tests/assets/std_env.hdr')
scn = synth.synth_sponza(tex_size=64, detail=0.05)
o = O.Oracle(); o.set_scene(scn)
res = Resources(scene=scn, hdr_rgb=env, camera=scn.camera, size=(16, 16))
pt = PathTracer(0); pt.onAttach(res)
rays = random_rays(400000, [-15, 0, -6], [15, 12, 6], seed=7)
h = o.trace_closest(rays, threads=64)
ok = h.view(np.int32)[:, 1] >= 0
p = rays[ok, 0:3] + rays[ok, 4:7] * h[ok, 0:1]
rng = np.random.default_rng(3)
d = rng.normal(size=p.shape); d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros((len(p), 8), np.float32); sec[:, 0:3] = p; sec[:, 4:7] = d; sec[:, 7] = 1e32
ref = o.trace_closest(sec, threads=64)
d_r = torch.from_numpy(sec).cuda(); d_h = torch.empty((len(sec), 6), dtype=torch.float32, device='cuda')
pt.trace_closest(d_r.data_ptr(), len(sec), d_h.data_ptr()); pt.synchronize()
got = d_h.cpu().numpy()
bad = ~((got.view(np.uint32) == ref.view(np.uint32)).all(axis=1))
print('secondary rays', len(sec), 'mismatch', bad.sum())
for i in np.where(bad)[0][:10]:
    print(sec[i].tolist(), 'ref', ref[i, 0], ref.view(np.int32)[i, 1:4], ref[i, 4:], 'got', got[i, 0], got.view(np.int32)[i, 1:4], got[i, 4:])
np.save('gpurun_out/bad_rays.npy', sec[bad]); np.save('gpurun_out/bad_ref.npy', ref[bad]); np.save('gpurun_out/bad_got.npy', got[bad])
# shadow
sec[:, 7] = 3.0
refs = o.trace_shadow(sec)
d_r = torch.from_numpy(sec).cuda(); d_t = torch.empty((len(sec), 3), dtype=torch.float32, device='cuda')
pt.trace_shadow(d_r.data_ptr(), len(sec), d_t.data_ptr()); pt.synchronize()
gs = d_t.cpu().numpy()
print('shadow mismatch', (gs != refs).any(axis=1).sum())
