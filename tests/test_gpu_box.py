"""-m gpu parity: CUDA path (through the C-ABI) vs the CPU oracle on BASELINE config 1."""
import numpy as np
import pytest

from conftest import rel_rmse

pytestmark = pytest.mark.gpu


def _attach(scene, env, size, **kw):
    from vk_gltf_renderer_b200.renderer import PathTracer, Resources
    res = Resources(scene=scene, hdr_rgb=env, camera=scene.camera, size=size)
    pt = PathTracer(0)
    for k, v in kw.items():
        setattr(pt, k, v)
    pt.onAttach(res)
    return pt, res


def test_environment_integral(box_scene, std_env, oracle_mod):
    pt, res = _attach(box_scene, std_env, (64, 64))
    o = oracle_mod.Oracle()
    assert o.set_environment(std_env) == pytest.approx(pt.hdr_integral, rel=0, abs=0)


def test_trace_parity_box(box_scene, std_env, oracle_mod):
    import torch
    from gpu_util import primary_rays, random_rays, to_dev
    pt, res = _attach(box_scene, std_env, (64, 64))
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    rays = np.concatenate([primary_rays(box_scene.camera, 128, 128), random_rays(50000, [-1, -1, -1], [1, 1, 1])])
    ref = o.trace_closest(rays)
    d_rays = to_dev(rays)
    d_hits = torch.empty((len(rays), 6), dtype=torch.float32, device="cuda")
    pt.trace_closest(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
    pt.synchronize()
    got = d_hits.cpu().numpy()
    # ids bit-exact, (t,u,v) bit-exact: same fma chain on both sides
    assert np.array_equal(got.view(np.uint32)[:, 1:4], ref.view(np.uint32)[:, 1:4])
    assert np.array_equal(got[:, [0, 4, 5]], ref[:, [0, 4, 5]])
    assert (ref.view(np.int32)[:, 1] >= 0).sum() > 1000


def test_render_parity_box_config1(box_scene, std_env, oracle_mod):
    """BASELINE config 1: Box.glb, 256x256, 16 frames x 1 spp, depth 4, std_env.hdr.
    Tolerance: per-pixel radiance <= 1e-3 relative RMSE (north_star); .w (solid flag mean) exact."""
    from vk_gltf_renderer_b200.renderer import render_headless, Resources
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    ref = oracle_mod.render(o, box_scene.camera, 256, 256, 16, max_depth=4)
    res = Resources(scene=box_scene, hdr_rgb=std_env, camera=box_scene.camera, size=(256, 256))
    pt, img = render_headless(res, 16, ptMaxDepth=4)
    assert img.shape == ref.shape
    assert np.isfinite(img).all()
    assert np.array_equal(img[..., 3], ref[..., 3])
    e = rel_rmse(img, ref)
    print("rel RMSE", e)
    assert e <= 1e-3
    st, so = pt.stats(), o.stats()
    assert st["closestRays"] == so["closestRays"] and st["shadowRays"] == so["shadowRays"]
    assert st["shadedHits"] == so["shadedHits"] and st["pathsStarted"] == so["paths"]
