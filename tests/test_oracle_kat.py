"""CPU: known-answer / property tests that pin the oracle (the reference holds no golden vectors for this path,
SURVEY.md section 8c — these analytic checks are what stands in for them)."""
import ctypes as C
import math

import numpy as np
import pytest


def _xxhash32(x, y, z):
    M = 0xFFFFFFFF
    P1, P2, P3, P4 = 2246822519, 3266489917, 668265263, 374761393
    rot = lambda v, r: ((v << r) | (v >> (32 - r))) & M
    h = (z + P4 + x * P2) & M
    h = (P3 * rot(h, 17)) & M
    h = (h + y * P2) & M
    h = (P3 * rot(h, 17)) & M
    h = (P1 * (h ^ (h >> 15))) & M
    h = (P2 * (h ^ (h >> 13))) & M
    return h ^ (h >> 16)


def _pcg(state):
    M = 0xFFFFFFFF
    prev = (state * 747796405 + 2891336453) & M
    word = (((prev >> ((prev >> 28) + 4)) ^ prev) * 277803737) & M
    return prev, ((word >> 22) ^ word) & M


def test_rng_contract(oracle_mod):
    """seed = xxhash32(x, y, frame) (Jarzynski-Olano), rand = PCG -> [0,1) via the 23-bit mantissa trick."""
    L = oracle_mod.lib()
    for x, y, z in ((0, 0, 0), (137, 221, 0), (1919, 1079, 255), (5, 7, 1 << 20)):
        assert L.oracle_xxhash32(x, y, z) == _xxhash32(x, y, z)
    s = C.c_uint32(_xxhash32(3, 4, 5))
    st = s.value
    for _ in range(100):
        got = L.oracle_rand(C.byref(s))
        st, r = _pcg(st)
        exp = np.frombuffer(np.uint32(0x3F800000 | (r >> 9)).tobytes(), np.float32)[0] - np.float32(1.0)
        assert got == exp and 0.0 <= got < 1.0 and s.value == st


def test_safe_offset_ray_bit_patterns(oracle_mod):
    """Waechter-Binder offset (pathtrace_functions.h.slang:151-167): integer ULP steps away from the origin,
    a float offset inside |p| < 1/32."""
    L = oracle_mod.lib()

    def off(p, n):
        p, n, o = np.array(p, np.float32), np.array(n, np.float32), np.zeros(3, np.float32)
        L.oracle_safe_offset_ray(p.ctypes.data, n.ctypes.data, o.ctypes.data)
        return o
    p = np.array([10.0, -3.0, 0.001], np.float32)
    o = off(p, [1.0, 1.0, 1.0])
    bits = lambda v: int(np.float32(v).view(np.int32))
    assert bits(o[0]) - bits(p[0]) == 256          # positive coordinate: +256 ulp
    assert bits(o[1]) - bits(p[1]) == -256         # negative coordinate: the int offset flips sign -> moves toward +n
    assert o[1] > p[1]
    assert o[2] == np.float32(0.001) + np.float32(1.0 / 65536.0)  # near the origin: float offset
    assert np.array_equal(off(p, [0, 0, 0]), p)


def test_alias_table_and_pdf_normalisation(oracle_mod, std_env):
    """sum over texels of pdf(alpha) * solid angle = 1; the alias table reproduces the importance distribution."""
    o = oracle_mod.Oracle()
    integral = o.set_environment(std_env)
    rgba, alias, q = o.get_environment()
    h, w = std_env.shape[:2]
    th = np.arange(h + 1, dtype=np.float64) * math.pi / h
    area = (np.cos(th[:-1]) - np.cos(th[1:])) * (2 * math.pi / w)
    # the integral is a sequential fp32 sum over 1.1M texels (std::accumulate with a float init in the external code): ~4e-4 drift
    assert float((rgba[..., 3].astype(np.float64) * area[:, None]).sum()) == pytest.approx(1.0, rel=1e-3)
    imp = (std_env.max(-1).astype(np.float64) * area[:, None]).reshape(-1)
    assert integral == pytest.approx(imp.sum(), rel=1e-3)
    # alias method: P(i) = (q_i + sum_{j: alias_j = i} (1 - q_j)) / N
    n = w * h
    prob = q.astype(np.float64).copy()
    np.add.at(prob, alias, 1.0 - q.astype(np.float64))
    prob /= n
    assert 0.5 * np.abs(prob - imp / imp.sum()).sum() < 1e-3   # total-variation distance (fp32 table)
    assert (q >= 0).all() and (q <= 1.0 + 1e-5).all()


def _single_tri_scene(double_sided=0):
    from vk_gltf_renderer_b200.scene import Scene
    s = Scene()
    m = s.add_material(pbrBaseColorFactor=[1, 1, 1, 1], pbrMetallicFactor=0.0, doubleSided=double_sided)
    p = s.add_primitive([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])
    s.add_node(p, m, [[2, 0, 0, 1], [0, 2, 0, 0], [0, 0, 2, 5], [0, 0, 0, 1]])  # scale 2, translate (1, 0, 5)
    return s


def test_single_triangle_hit_bary_t_and_culling(oracle_mod):
    """world triangle (1,0,5) (3,0,5) (1,2,5), CCW seen from -z... the front face (normal +z) is seen from z > 5."""
    o = oracle_mod.Oracle()
    o.set_scene(_single_tri_scene())
    rays = np.array([[1.5, 0.5, 9, 0, 0, 0, -1, 1e32],      # front side: hit at t=4, u=.25, v=.25
                     [1.5, 0.5, 1, 0, 0, 0, 1, 1e32],       # back side, single-sided: culled
                     [1.5, 0.5, 9, 0, 0, 0, -1, 3.9],       # tmax before the plane
                     [5.0, 5.0, 9, 0, 0, 0, -1, 1e32]], np.float32)
    h = o.trace_closest(rays)
    ids = h.view(np.int32)
    assert h[0, 0] == pytest.approx(4.0) and ids[0, 1:4].tolist() == [0, 0, 0] and h[0, 4:].tolist() == pytest.approx([0.25, 0.25])
    assert ids[1, 1] == -1 and h[1, 0] == np.float32(1e32)
    assert ids[2, 1] == -1 and ids[3, 1] == -1
    t = o.trace_shadow(rays)          # shadow rays never cull (raytracer_interface.h.slang:147)
    assert t[0].tolist() == [0, 0, 0] and t[1].tolist() == [0, 0, 0] and t[2].tolist() == [1, 1, 1]
    o2 = oracle_mod.Oracle()
    o2.set_scene(_single_tri_scene(double_sided=1))
    assert o2.trace_closest(rays).view(np.int32)[1, 1] == 0   # doubleSided disables culling (gltf_scene_rtx.cpp:289-292)


def test_furnace_lambert_and_energy(oracle_mod, box_scene):
    """White furnace (constant environment = 1): a pure Lambertian surface returns exactly its albedo in
    expectation (NEE + BSDF sampling + MIS sum to one estimator); rough dielectric / metal never gain energy."""
    import copy
    from vk_gltf_renderer_b200 import scene as scn_mod
    import os
    s = scn_mod.load_gltf(os.path.join(os.path.dirname(__file__), "assets", "Box.glb"))
    s.materials[0].specularFactor = 0.0
    o = oracle_mod.Oracle()
    o.set_scene(s)
    o.set_environment(np.ones((16, 32, 3), np.float32))
    img = oracle_mod.render(o, s.camera, 16, 16, 1024, max_depth=4)
    face = img[5:11, 5:11, :3].mean((0, 1))
    assert face[0] == pytest.approx(0.8, rel=6e-3) and face[1] == 0.0 and face[2] == 0.0
    assert img[0, 0].tolist() == [1.0, 1.0, 1.0, 0.0]     # primary miss: env radiance, solid flag 0
    assert img[8, 8, 3] == 1.0
    s.materials[0].specularFactor = 1.0
    s.materials[0].pbrBaseColorFactor[:] = [1, 1, 1, 1]
    s.materials[0].pbrMetallicFactor = 1.0
    s.materials[0].pbrRoughnessFactor = 0.3
    o.set_scene(s)
    img = oracle_mod.render(o, s.camera, 16, 16, 512, max_depth=8)
    v = img[5:11, 5:11, 0].mean()
    assert 0.93 < v <= 1.0 + 5e-3


def test_bsdf_sample_eval_consistency(oracle_mod):
    """For every sampled direction, evaluating the same lobe (same xi.z) must return the sampler's pdf, and
    bsdf_over_pdf must equal (diffuse+glossy)/pdf: the two halves of the external BSDF agree with each other."""
    from vk_gltf_renderer_b200 import bsdf_io
    o = oracle_mod.Oracle()
    rec = bsdf_io.random_records(60000, seed=7)
    rec[:, 3:5] = np.maximum(rec[:, 3:5], 1e-3)  # alpha >= 1e-3: below that the GGX peak (pdf ~ 1e8) is fp32-noisy
    smp = o.bsdf_sample(rec)
    live = smp[:, 7] != 0
    rec2 = rec.copy()
    rec2[:, 42:45] = smp[:, 0:3]
    ev = o.bsdf_eval(rec2)
    assert live.mean() > 0.4
    pdf_s, pdf_e = smp[live, 6], ev[live, 6]
    ok = np.isclose(pdf_s, pdf_e, rtol=2e-3, atol=1e-6)
    assert ok.mean() > 0.995
    w = (ev[live, 0:3] + ev[live, 3:6]) / np.maximum(pdf_e[:, None], 1e-30)
    ok2 = np.isclose(w, smp[live, 3:6], rtol=5e-3, atol=1e-4).all(1)
    assert (ok & ok2).mean() > 0.99
    assert (smp[live, 3:6] <= 1.0 + 1e-3).all() and (smp[:, 3:6] >= 0).all()   # single-scatter lobes never gain energy
    k2 = smp[live, 0:3]
    assert np.allclose(np.linalg.norm(k2, axis=1), 1.0, atol=1e-4)


def test_texture_sampler_matches_numpy_reference(oracle_mod):
    """bilinear at texel centres returns the texel; REPEAT wraps; the gradient picks the mip level log2(g*size);
    sRGB textures decode before filtering."""
    from vk_gltf_renderer_b200.scene import Scene
    rng = np.random.default_rng(5)
    tex = rng.integers(0, 256, (16, 16, 4), dtype=np.uint8)
    s = Scene()
    s.add_material()
    s.add_primitive([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]])
    s.add_node(0, 0)
    s.add_texture(tex, srgb=False)
    s.add_texture(tex, srgb=True)
    o = oracle_mod.Oracle()
    o.set_scene(s)
    for (x, y) in ((0, 0), (5, 9), (15, 15)):
        got = o.sample_texture(0, (x + 0.5) / 16, (y + 0.5) / 16)
        assert np.allclose(got, tex[y, x] / 255.0, atol=1e-6)
    a = o.sample_texture(0, 1.0 / 16, 0.5 / 16)            # halfway between texel 0 and 1 of row 0
    assert np.allclose(a, (tex[0, 0] / 255.0 + tex[0, 1] / 255.0) / 2, atol=1e-6)
    assert np.allclose(o.sample_texture(0, 1.0 + 5.5 / 16, -1.0 + 2.5 / 16), tex[2, 5] / 255.0, atol=1e-6)
    lin = lambda c: np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    got = o.sample_texture(1, 3.5 / 16, 3.5 / 16)
    assert np.allclose(got[:3], lin(tex[3, 3, :3] / 255.0), atol=1e-6) and got[3] == pytest.approx(tex[3, 3, 3] / 255.0)
    # g * size = 16 -> lambda = 4 = coarsest level (1x1): the average colour, independent of uv
    top = o.sample_texture(0, 0.3, 0.7, g=1.0)
    assert np.allclose(top, o.sample_texture(0, 0.9, 0.1, g=1.0), atol=1e-6)
    assert np.allclose(top[:3], tex[..., :3].mean((0, 1)) / 255.0, atol=0.02)
    # g * size = 2 -> lambda = 1: exactly level 1 (8x8), whose texel 0 is the 2x2 box average (8-bit requantised)
    l1 = o.sample_texture(0, 0.5 / 8, 0.5 / 8, g=2.0 / 16)
    assert np.allclose(l1, np.floor(tex[0:2, 0:2].reshape(4, 4).mean(0) + 0.5) / 255.0, atol=1e-6)


def test_accumulation_is_running_mean(oracle_mod, box_scene, std_env):
    """frame 0 overwrites, later frames are (old*total + new*n)/(total+n) (gltf_pathtrace.slang:619-630)."""
    from vk_gltf_renderer_b200 import camera as cm
    o = oracle_mod.Oracle()
    o.set_scene(box_scene)
    o.set_environment(std_env)
    cam = box_scene.camera
    fi = cm.make_frame_info(cam, 32, 32)
    frames = []
    for f in range(3):
        a = np.full((32, 32, 4), 123.0, np.float32)
        pc = cm.make_push_constant(cam, 32, frame_count=0 if f == 0 else f, total_samples=0, max_depth=3)
        pc.frameCount = f
        pc.flags = 4  # first-frame flag: overwrite -> isolates each frame's own estimate
        o.render_frame(fi, pc, a)
        frames.append(a)
    acc = oracle_mod.render(o, cam, 32, 32, 3, max_depth=3)
    exp = frames[0]
    exp = (exp * np.float32(1) + frames[1] * np.float32(1)) / np.float32(2)
    exp = (exp * np.float32(2) + frames[2] * np.float32(1)) / np.float32(3)
    assert np.array_equal(acc, exp)


def test_deep_layer_scenes_are_deterministic_and_stochastic(std_env, oracle_mod):
    """The 14-layer any-hit stress scenes (used by the GPU continuation tests): the oracle is deterministic per
    (pixel, frame) seed, frames differ (stochastic alpha), and MASK coverage lets some light through."""
    from vk_gltf_renderer_b200 import synth
    for kw in ({}, {"blend": True}, {"tinted": True}):
        scn = synth.synth_layers(**kw)
        o = oracle_mod.Oracle()
        o.set_scene(scn)
        o.set_environment(std_env)
        a = oracle_mod.render(o, scn.camera, 64, 48, 2, max_depth=4)
        b = oracle_mod.render(o, scn.camera, 64, 48, 2, max_depth=4)
        c = oracle_mod.render(o, scn.camera, 64, 48, 3, max_depth=4)
        assert np.array_equal(a, b) and np.isfinite(a).all()
        assert not np.array_equal(a, c)
        assert a[..., :3].mean() > 0.01
