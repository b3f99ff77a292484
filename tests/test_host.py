"""CPU: host-side mirror of the reference's data preparation (loader, camera, HDR, synthetic scenes, tiling)."""
import math
import os

import numpy as np
import pytest


def test_box_glb_matches_reference_facts(box_scene):
    """SURVEY.md Appendix B 'Box.glb facts': 24 vertices, 12 triangles, red dielectric, camera from node extras."""
    s = box_scene
    assert len(s.render_nodes) == 1 and len(s.render_prims) == 1 and len(s.materials) == 1
    p = s.render_prims[0]
    assert p["positions"].shape == (24, 3) and p["indices"].shape == (12, 3) and p["indices"].dtype == np.uint32
    assert p["normals"] is not None and p["uv0"] is None and p["tangents"] is None
    m = s.materials[0]
    assert list(m.pbrBaseColorFactor) == pytest.approx([0.8, 0.0, 0.0, 1.0])
    assert m.pbrMetallicFactor == 0.0 and m.pbrRoughnessFactor == 1.0 and m.alphaMode == 0 and m.doubleSided == 0
    assert m.specularFactor == 1.0 and m.ior == 1.5 and m.pbrBaseColorTexture == 0
    c = s.camera
    assert c.eye.tolist() == pytest.approx([0, 0, 2.0905852]) and c.center.tolist() == [0, 0, 0]
    assert c.yfov == pytest.approx(0.7853982)
    assert len(s.texture_infos) == 1 and s.texture_infos[0].index == -1  # slot 0 = "no texture"


def test_shader_ball_loads(shader_ball_scene):
    s = shader_ball_scene
    assert s.num_triangles() == 9450 and s.materials[0].doubleSided == 1
    assert s.materials[0].pbrRoughnessFactor == pytest.approx(0.6)
    assert s.render_prims[0]["uv0"] is not None


def test_hdr_loader(std_env):
    assert std_env.shape == (750, 1500, 3) and std_env.dtype == np.float32
    assert np.isfinite(std_env).all() and std_env.min() >= 0 and 5 < std_env.max() < 100


def test_camera_matrices_roundtrip(box_scene):
    from vk_gltf_renderer_b200 import camera as cm
    fi = cm.make_frame_info(box_scene.camera, 256, 256)
    glm = lambda a: np.array(a[:], np.float64).reshape(4, 4).T
    v, vi = glm(fi.viewMatrix), glm(fi.viewInv)
    assert np.allclose(v @ vi, np.eye(4), atol=1e-5)
    assert vi[:3, 3].tolist() == pytest.approx([0, 0, 2.0905852], abs=1e-5)
    pc = cm.make_push_constant(box_scene.camera, 256, frame_count=0, total_samples=0)
    assert pc.flags == 4 and pc.pixelAngle == pytest.approx(2 * math.tan(0.7853982 / 2) / 256, rel=1e-6)
    assert pc.focalDistance == pytest.approx(2.0905852, rel=1e-6)
    assert cm.make_push_constant(box_scene.camera, 256, frame_count=3, total_samples=3).flags == 0


def test_texture_transform_matches_reference_quirk():
    from vk_gltf_renderer_b200.scene import _tex_transform
    t = _tex_transform({"extensions": {"KHR_texture_transform": {"offset": [0.1, 0.2], "rotation": 0.5, "scale": [2, 3]}}})
    c, s = math.cos(0.5), math.sin(0.5)
    assert t == pytest.approx((2 * c, -3 * s, 2 * s, 3 * c, 0.1, 0.2))


def test_synth_sponza_is_deterministic_and_sized():
    from vk_gltf_renderer_b200 import synth
    a = synth.synth_sponza(tex_size=32, detail=0.02)
    b = synth.synth_sponza(tex_size=32, detail=0.02)
    assert a.num_triangles() == b.num_triangles() and len(a.materials) == 25
    for pa, pb in zip(a.render_prims, b.render_prims):
        assert np.array_equal(pa["positions"], pb["positions"])
    assert np.array_equal(a.textures[0]["rgba8"], b.textures[0]["rgba8"])
    st = synth.scene_from_state(synth.scene_state(a))
    assert st.num_triangles() == a.num_triangles() and bytes(st.materials[3]) == bytes(a.materials[3])


def test_tiling_partition_covers_the_frame():
    from vk_gltf_renderer_b200 import tiling
    for h, w in ((1080, 8), (7, 4), (1, 2), (2160, 8)):
        rows = [tiling.partition_rows(h, w, r) for r in range(w)]
        assert sum(n for _, n in rows) == h
        y = 0
        for y0, n in rows:
            if n:
                assert y0 == y
                y += n


def test_scene_blob_and_cpp_host_fail_loudly_without_gpu(tmp_path, box_scene, std_env):
    """Scene.save_blob writes what host/b200pt_host.cpp loads; without a CUDA device the C++ headless driver exits
    non-zero with the library's error (no CPU fallback anywhere).  On a GPU box the same command succeeds."""
    import os
    import struct
    import subprocess
    from vk_gltf_renderer_b200 import _lib
    blob = str(tmp_path / "box.b2sc")
    box_scene.save_blob(blob, std_env)
    raw = open(blob, "rb").read()
    assert raw[:4] == b"B2SC"
    ver, nn, npr, nm, nti, nt, nl = struct.unpack("<7I", raw[4:32])
    assert (ver, nn, npr, nm) == (1, len(box_scene.render_nodes), len(box_scene.render_prims), len(box_scene.materials))
    assert len(raw) > std_env.size * 4
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "b200pt_headless")
    if not os.path.exists(exe):
        pytest.skip("b200pt_headless not built")
    out = subprocess.run([exe, "--scene", blob, "--size", "32", "32", "--frames", "1"], capture_output=True, text=True, timeout=300)
    if out.returncode != 0:
        assert "b200pt_create failed" in out.stderr or "CUDA" in out.stderr
    else:
        assert "HEADLESS_SUMMARY" in out.stdout
    bad = subprocess.run([exe, "--scene", str(tmp_path / "missing.b2sc")], capture_output=True, text=True, timeout=60)
    assert bad.returncode != 0 and "cannot open" in bad.stderr


def test_interleave_band_choices_and_bench_core_count():
    """Band height: the largest <= 8 that tiles height / world; None when the height does not split evenly.
    bench.usable_cores() never reports more than the scheduler affinity and at least one core."""
    import sys
    from vk_gltf_renderer_b200 import tiling
    assert tiling.interleave_band(1080, 1) is None
    assert tiling.interleave_band(1080, 2) == 6 and tiling.interleave_band(1080, 8) == 5 and tiling.interleave_band(2160, 8) == 6
    assert tiling.interleave_band(37, 2) is None
    for world in (2, 4, 8):
        band = tiling.interleave_band(1080, world)
        rows = [tiling.interleaved_rows(1080, world, r, band) for r in range(world)]
        flat = sorted(y for rr in rows for y in rr)
        assert flat == list(range(1080)) and all(len(rr) == 1080 // world for rr in rows)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_adaptive_sampling_controller_follows_the_reference_rules():
    """PathTracer::updateAdaptiveSampling (src/renderer_pathtracer.cpp:1326-1374): reset to 1 at frame 0, untouched before frame 5,
    +1 below 80 % of the target frame time, -1 above 110 %, clamped to [1, 100]; off = never touched."""
    from vk_gltf_renderer_b200.renderer import PathTracer, Resources
    pt = PathTracer.__new__(PathTracer)
    pt.ptAdaptiveSampling, pt.ptPerformanceTarget, pt.ptSamples, pt.last_frame_gpu_ms = True, 1, 7, None
    res = Resources()
    res.frameCount = 0
    pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 1
    pt.last_frame_gpu_ms = 5.0
    for f in range(1, 5):
        res.frameCount = f
        pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 1                       # first frames: hands off
    for f in range(5, 9):
        res.frameCount = f
        pt.updateAdaptiveSampling(res)             # 5 ms << 0.8 * 33.3 ms: headroom
    assert pt.ptSamples == 5
    pt.last_frame_gpu_ms = 30.0                    # inside [0.8, 1.1] x target: hold
    pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 5
    pt.last_frame_gpu_ms = 40.0                    # > 1.1 x 33.3 ms: back off
    for _ in range(10):
        pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 1
    pt.ptPerformanceTarget, pt.last_frame_gpu_ms, pt.ptSamples = 3, 1.0, 99
    pt.updateAdaptiveSampling(res)
    pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 100 and pt.getTargetFrameTimeMs() == 100.0
    pt.ptAdaptiveSampling, pt.ptSamples = False, 4
    pt.updateAdaptiveSampling(res)
    assert pt.ptSamples == 4


def test_cpp_blob_writer_round_trip(tmp_path, box_scene, std_env):
    """host/b2sc_writer.hpp (the scene-blob writer a maintainer drops into the reference loader: it serialises the C-ABI
    b200pt_scene_desc): a blob written by the Python loader, loaded by the C++ SceneData and written back through
    the header is byte-identical -- for Box.glb with its environment and for a textured scene that carries opacity micromaps, lights,
    TEXCOORD_1, tangents and vertex colours."""
    import os
    import subprocess
    from vk_gltf_renderer_b200 import _lib, omm, synth
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "b200pt_headless")
    if not os.path.exists(exe):
        pytest.skip("b200pt_headless not built")
    zoo = synth.synth_material_zoo()
    zoo.lights = synth.synth_lit().lights
    atrium = synth.synth_sponza(tex_size=64, detail=0.05)
    omm.bake_opacity_micromaps(atrium, level=3)
    assert atrium.prim_omms
    for name, scn, env in (("box", box_scene, std_env), ("zoo", zoo, None), ("atrium", atrium, None)):
        a, b = str(tmp_path / (name + "_a.b2sc")), str(tmp_path / (name + "_b.b2sc"))
        scn.save_blob(a, env)
        out = subprocess.run([exe, "--scene", a, "--writeBlob", b], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        ra, rb = open(a, "rb").read(), open(b, "rb").read()
        assert len(ra) == len(rb) and ra == rb, name
