"""Shadow-catcher plane in the oracle, without a GPU (handleShadowCatcher, shaders/pathtrace_functions.h.slang:499-554): known answers on
Box.glb -- a plane point whose light sample is unblocked shows exactly the environment behind it (first hit: lastSamplePdf is DIRAC, MIS
weight 1), blocked ones are darker, never brighter, shadowCatcherDarkness darkens further, and the catcher counts as a solid hit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_oracle_shadow_catcher_known_answers(box_scene, std_env, oracle_mod):
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    cam = box_scene.camera
    W, H = 96, 64
    lo = float(min(p["positions"][:, 1].min() for p in box_scene.render_prims))
    kw = dict(infinite_plane=True, plane_distance=lo - 0.6, plane_color=(0.5, 0.5, 0.5), plane_roughness=0.5)
    bare = oracle_mod.render(o, cam, W, H, 1, max_depth=4)
    solid = oracle_mod.render(o, cam, W, H, 1, max_depth=4, shadow_catcher=False, **kw)
    catch = oracle_mod.render(o, cam, W, H, 1, max_depth=4, shadow_catcher=True, **kw)
    dark = oracle_mod.render(o, cam, W, H, 1, max_depth=4, shadow_catcher=True, catcher_darkness=0.8, **kw)
    assert np.isfinite(catch).all() and np.isfinite(dark).all()
    plane = (bare[..., 3] == 0) & (solid[..., 3] == 1)          # primary ray passes the box and meets the plane
    assert plane.sum() > 300
    same = np.all(catch[..., :3] == bare[..., :3], -1)
    assert same[plane].mean() > 0.5                              # lit plane points: the environment, bit for bit
    # ... and they are solid hits (pt.solid stays true).  A SHADOWED point whose continuation ray escapes is not: the continuation
    # still runs at surfaceDepth 0, so tryPrimaryMissBackplate clears pt.solid (pathtrace_functions.h.slang:946) -- the reference's
    # own behaviour, followed
    assert (catch[..., 3][plane & same] > 0).all()
    assert (same | ~plane).mean() > 0.5
    shadowed = plane & ~same
    assert shadowed.sum() > 10
    lum = lambda a: a[..., :3].sum(-1)
    # one sample per pixel, first hit: radiance = env * shadowFactor (- darkening) + continuation light; never above the plain env
    # by more than the continuation can add, and the darkened variant is never brighter than the plain catcher
    assert lum(dark)[shadowed].sum() < lum(catch)[shadowed].sum() <= lum(bare)[shadowed].sum() * 1.5
    assert np.array_equal(dark[..., :3][plane & same], catch[..., :3][plane & same])
    assert not np.array_equal(catch[..., :3], solid[..., :3])
