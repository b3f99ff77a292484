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


def test_bsdf_sample_simple_known_answers(oracle_mod):
    """bsdfSampleSimple as restated (nvshaders, external): the sampled direction lies in the upper hemisphere, a white diffuse
    dielectric returns (almost) all energy and never creates any, a metal returns a little less than its base colour (single-scatter
    GGX; Schlick lifts the dark channels), roughness -> 0 gives the mirror direction, and the Monte-Carlo mean of bsdf_over_pdf equals
    the quadrature of value over the hemisphere (the sampling is unbiased for its own evaluation)."""
    o = oracle_mod.Oracle()
    n = 40000
    rng = np.random.default_rng(5)

    def records(base, metallic, alpha, cos_in):
        r = np.zeros((n, 48), np.float32)
        r[:, 0:3] = base
        r[:, 3:5] = alpha
        r[:, 5] = metallic
        r[:, 6:9], r[:, 9:12], r[:, 12:15], r[:, 15:18] = (0, 0, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1)
        r[:, 18], r[:, 19], r[:, 20], r[:, 21:24] = 1.0, 1.5, 1.0, 1.0
        s = np.sqrt(1 - cos_in * cos_in)
        r[:, 39:42] = (s, 0.0, cos_in)
        r[:, 45:48] = rng.random((n, 3))
        return r
    white = o.bsdf_sample_simple(records((1, 1, 1), 0.0, 0.5, 0.8))
    ok = white[:, 7] != 0
    assert ok.mean() > 0.9 and (white[ok, 2] > 0).all()
    assert np.allclose(np.linalg.norm(white[ok, 0:3], axis=1), 1.0, atol=1e-4)
    mean = (white[:, 3:6] * ok[:, None]).mean(0)
    assert (mean < 1.02).all() and (mean > 0.85).all(), mean           # energy conserving, nearly white
    metal = o.bsdf_sample_simple(records((0.9, 0.6, 0.3), 1.0, 0.25, 0.7))
    okm = metal[:, 7] != 0
    mm = (metal[:, 3:6] * okm[:, None]).mean(0)
    ratio = mm / np.float32([0.9, 0.6, 0.3])   # single-scatter GGX loses some energy to masking; Schlick lifts the dark channels a little
    assert (mm < 1.0).all() and mm[0] > mm[1] > mm[2] and (ratio > 0.8).all() and (ratio < 1.02).all() and ratio[2] >= ratio[0], (mm, ratio)
    assert (metal[okm, 7] == 10).all()                                 # metallic 1: only the glossy lobe (GLOSSY | REFLECTION)
    mirror = o.bsdf_sample_simple(records((1, 1, 1), 1.0, 1e-4, 0.6))
    okr = mirror[:, 7] != 0
    dev = np.abs(mirror[okr, 0:3] - np.float32([-0.8, 0.0, 0.6])).max(1)   # GGX has a long tail: nearly all samples within 2e-3, all close
    assert okr.mean() > 0.99 and (dev < 2e-3).mean() > 0.98 and dev.max() < 0.1
