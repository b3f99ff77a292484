"""Dump world-space triangles + a mixed ray set (primary + incoherent secondary) for csrc/tools/bvh_stats."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from gpu_util import primary_rays
from oracle import oracle as O
import bench
class A: pass
a = A(); a.tex = 64; a.detail = 1.0
from vk_gltf_renderer_b200 import synth
scn = synth.synth_sponza(seed=1234, tex_size=64, detail=1.0)
tr = []
for rn in scn.render_nodes:
    p = scn.render_prims[rn['renderPrimID']]
    m = rn['objectToWorld'].reshape(4, 4).T.astype(np.float64)
    w = (p['positions'].astype(np.float64) @ m[:3, :3].T + m[:3, 3]).astype(np.float32)
    t = w[p['indices']]
    tr.append(np.concatenate([t[:, 0], t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]], 1))
tr = np.concatenate(tr).astype(np.float32)
o = O.Oracle(); o.set_scene(scn)
pr = primary_rays(scn.camera, 480, 270)
h = o.trace_closest(pr, threads=8)
ok = h.view(np.int32)[:, 1] >= 0
pts = pr[ok, 0:3] + pr[ok, 4:7] * (h[ok, 0:1] * 0.9999)
rng = np.random.default_rng(1)
d = rng.normal(size=pts.shape); d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros((len(pts), 8), np.float32); sec[:, :3] = pts; sec[:, 4:7] = d; sec[:, 7] = 1e32
rays = np.concatenate([pr, sec]).astype(np.float32)
with open('/tmp/bvh_dump.bin', 'wb') as f:
    f.write(np.array([len(tr), len(rays)], np.uint32).tobytes()); f.write(tr.tobytes()); f.write(rays.tobytes())
print(len(tr), len(rays))
