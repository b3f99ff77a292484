"""Opacity micromaps on the CPU side: the micro-triangle order, the baker's conservativeness, and the checker's semantics
(reference: src/gltf_scene_omm.cpp consumes EXT_mesh_opacity_micromap; docs/RENDERING_ARCHITECTURE.md:65-78 says what the traversal
does with it: OPAQUE micro-triangles commit, TRANSPARENT ones are culled, UNKNOWN ones run the any-hit alpha logic)."""
import numpy as np
import pytest

from vk_gltf_renderer_b200 import abi, camera, omm, synth


def test_bary2index_is_a_hierarchical_bijection():
    """What pins the restated VK_EXT_opacity_micromap bary2index: per level a bijection between micro-triangles and [0, 4^level),
    index >> 2 = the parent's index one level up (the 'bird curve' is hierarchical), level 1 = corner w, centre, corner u, corner v."""
    assert [int(omm.bary2index(np.float32(u), np.float32(v), 1)) for u, v in ((1 / 6, 1 / 6), (1 / 3, 1 / 3), (2 / 3, 1 / 6), (1 / 6, 2 / 3))] == [0, 1, 2, 3]
    for level in range(0, 7):
        n = 1 << level
        cents = []
        for i in range(n):
            for j in range(n - i):
                cents.append(((i + 1 / 3) / n, (j + 1 / 3) / n))
                if i + j < n - 1:
                    cents.append(((i + 2 / 3) / n, (j + 2 / 3) / n))
        c = np.asarray(cents, np.float32)
        idx = omm.bary2index(c[:, 0], c[:, 1], level)
        assert sorted(idx.tolist()) == list(range(4 ** level))
        if level:
            assert np.array_equal(idx >> 2, omm.bary2index(c[:, 0], c[:, 1], level - 1))
    # every point of a micro-triangle maps to its index: random points against the corner table
    rng = np.random.default_rng(5)
    for level in (2, 5):
        corners = omm.micro_triangle_corners(level)
        k = rng.integers(0, 4 ** level, 4000)
        w = rng.dirichlet([1, 1, 1], 4000) * 0.98 + 0.02 / 3     # strictly inside
        p = (corners[k] * w[:, :, None]).sum(1)
        assert np.array_equal(omm.bary2index(p[:, 0].astype(np.float32), p[:, 1].astype(np.float32), level), k.astype(np.uint32))


def test_pack_states_layout():
    st = np.array([[1, 0, 3, 2, 1, 1, 0, 0]], np.uint8)
    assert omm.pack_states(st, abi.OMM_FORMAT_4_STATE).tolist() == [[1 | (0 << 2) | (3 << 4) | (2 << 6), 1 | (1 << 2)]]
    assert omm.pack_states(st & 1, abi.OMM_FORMAT_2_STATE).tolist() == [[0b00110101]]


def _foliage_scene():
    return synth.synth_sponza(tex_size=256, detail=0.05)


def test_baker_is_conservative():
    """every micro-triangle the baker calls OPAQUE / TRANSPARENT agrees with the any-hit evaluation (the oracle's getOpacity through
    sample_texture semantics: level-0 texel, MASK cutoff) at random points inside it"""
    scn = _foliage_scene()
    st = omm.bake_opacity_micromaps(scn, level=4)
    assert st["triangles"] > 500 and 0.2 < st["unknown"] / st["micro"] < 0.7 and st["opaque"] > 0 and st["transparent"] > 0
    mm = scn.micromaps[0]
    rng = np.random.default_rng(3)
    checked = 0
    for po in scn.prim_omms[:3]:
        prim = scn.render_prims[po["renderPrimID"]]
        node = next(rn for rn in scn.render_nodes if rn["renderPrimID"] == po["renderPrimID"])
        m = scn.materials[node["materialID"]]
        tex = scn.textures[scn.texture_infos[m.pbrBaseColorTexture].index]
        a8 = tex["rgba8"][..., 3]
        H, W = a8.shape
        for t in rng.integers(0, len(prim["indices"]), 60):
            idx = po["indices"][t]
            if idx < 0:
                continue
            rec = mm["triangles"][idx]
            data = mm["data"][rec["dataOffset"]:]
            uv = prim["uv0"][prim["indices"][t]].astype(np.float32)
            b = rng.dirichlet([1, 1, 1], 200).astype(np.float32)
            p = b[:, 0:1] * uv[0] + b[:, 1:2] * uv[1] + b[:, 2:3] * uv[2]
            k = omm.bary2index(b[:, 1], b[:, 2], int(rec["subdivisionLevel"]))
            state = (data[k >> 2] >> ((k & 3) * 2)) & 3
            x = np.floor(p[:, 0] * np.float32(W)).astype(int) % W
            y = np.floor(p[:, 1] * np.float32(H)).astype(int) % H
            opaque = a8[y, x].astype(np.float32) / 255.0 * m.pbrBaseColorFactor[3] >= m.alphaCutoff
            assert np.all(opaque[state == 1]) and not np.any(opaque[state == 0])
            checked += int((state < 2).sum())
    assert checked > 2000


def test_oracle_with_micromaps_same_surfaces_fewer_any_hit_draws(std_env, oracle_mod):
    """The micromap only removes rand() draws: MASK decisions are deterministic, so the FIRST hit of every camera ray is the same
    surface with and without it; the images are two equally valid estimates (means agree within noise); ray-level: the committed
    hit of every ray is identical, only the seeds advance less."""
    from gpu_util import random_rays
    scn = _foliage_scene()
    o0 = oracle_mod.Oracle()
    o0.set_scene(scn)
    o0.set_environment(std_env)
    rays = random_rays(20000, [-15, 0, -6], [15, 12, 6])
    seeds = ((np.arange(len(rays), dtype=np.uint64) * 2654435761) % (2 ** 32)).astype(np.uint32)
    s0 = seeds.copy()
    h0 = o0.trace_closest(rays, s0)
    rays_s = rays.copy()
    rays_s[:, 7] = 4.0
    ss0 = seeds.copy()
    t0 = o0.trace_shadow(rays_s, ss0)
    img0 = oracle_mod.render(o0, scn.camera, 96, 54, 24, max_depth=4)

    st = omm.bake_opacity_micromaps(scn, level=4)
    o1 = oracle_mod.Oracle()
    o1.set_scene(scn)
    o1.set_environment(std_env)
    s1 = seeds.copy()
    h1 = o1.trace_closest(rays, s1)
    ss1 = seeds.copy()
    t1 = o1.trace_shadow(rays_s, ss1)
    # same committed hits; a transparent MASK texel is accepted only when rand() returns exactly 0 (2^-24), not in 20k rays
    assert np.array_equal(h0.view(np.uint32), h1.view(np.uint32))
    assert np.array_equal(t0, t1)       # MASK foliage: the shadow transmission is 0 or 1 either way
    drew0, drew1 = (s0 != seeds).sum(), (s1 != seeds).sum()
    assert drew1 < 0.6 * drew0 and drew0 > 500
    img1 = oracle_mod.render(o1, scn.camera, 96, 54, 24, max_depth=4)
    assert np.array_equal(img0[..., 3] > 0, img1[..., 3] > 0)
    m0, m1 = img0[..., :3].mean(), img1[..., :3].mean()
    assert abs(m0 - m1) / m0 < 0.03
    assert not np.array_equal(img0, img1)   # different random streams behind foliage


def test_special_indices_and_two_state_format(oracle_mod, std_env):
    """FULLY_OPAQUE / FULLY_TRANSPARENT special indices and the 1-bit format: a quad whose two triangles are forced either way"""
    scn = synth.synth_layers(layers=1, tex_size=32)
    o = oracle_mod.Oracle()
    pid = next(rn["renderPrimID"] for rn in scn.render_nodes if scn.materials[rn["materialID"]].alphaMode == 1)
    ntri = len(scn.render_prims[pid]["indices"])
    rays = np.zeros((4000, 8), np.float32)
    rng = np.random.default_rng(2)
    lo, hi = scn.bounds()
    rays[:, 0:3] = rng.uniform(lo - 0.5, hi + 0.5, (4000, 3))
    d = rng.normal(size=(4000, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 7] = 1e30
    seeds = np.arange(4000, dtype=np.uint32)
    o.set_scene(scn)
    base = o.trace_closest(rays, seeds.copy())
    out = {}
    for name, special in (("opaque", abi.OMM_INDEX_FULLY_OPAQUE), ("transparent", abi.OMM_INDEX_FULLY_TRANSPARENT)):
        scn.micromaps = [dict(data=np.zeros(1, np.uint8), triangles=np.zeros(0, abi.MICROMAP_TRIANGLE_DTYPE))]
        scn.prim_omms = [dict(renderPrimID=pid, micromap=0, baseTriangle=0, indices=np.full(ntri, special, np.int32))]
        s = seeds.copy()
        o.set_scene(scn)
        out[name] = (o.trace_closest(rays, s), s)
        assert np.array_equal(s, seeds)                      # nothing left for the any-hit path on that primitive
    bits = lambda a: a.view(np.uint32)
    assert not np.array_equal(bits(out["opaque"][0]), bits(out["transparent"][0])) and not np.array_equal(bits(out["opaque"][0]), bits(base))
    # 2-state, level 1: micro-triangle 1 (the centre) opaque, the corners transparent
    scn.micromaps = [dict(data=np.array([0b0010], np.uint8), triangles=np.array([(0, 1, abi.OMM_FORMAT_2_STATE)], abi.MICROMAP_TRIANGLE_DTYPE))]
    scn.prim_omms = [dict(renderPrimID=pid, micromap=0, baseTriangle=0, indices=np.zeros(ntri, np.int32))]
    o.set_scene(scn)
    s = seeds.copy()
    h = o.trace_closest(rays, s)
    on_prim = (h[:, 0] < 1e30) & (h.view(np.uint32)[:, 2] == pid)
    assert np.array_equal(s, seeds) and on_prim.sum() > 20
    u, v = h[on_prim, 4], h[on_prim, 5]
    assert np.all((u <= 0.5 + 1e-6) & (v <= 0.5 + 1e-6) & (u + v >= 0.5 - 1e-6))   # only the centre micro-triangle is ever hit


def test_loader_reads_ext_mesh_opacity_micromap(tmp_path, oracle_mod):
    """A .gltf carrying EXT_mesh_opacity_micromap (what src/gltf_scene_omm.cpp:140-391 parses): root micromaps[] with data /
    triangles bufferViews (one with a byteStride) and the per-primitive micromap / micromapBaseTriangle / micromapIndices (uint16:
    0xFFFF = FULLY_TRANSPARENT, like a VkIndexType).  The loaded scene drives the oracle: triangle 0 of the quad is culled, on
    triangle 1 only the centre micro-triangle (level 1, 1-bit format) stops a ray."""
    import base64
    import json
    from vk_gltf_renderer_b200 import scene as scn_mod
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tri_recs = np.zeros(2, np.dtype([("dataOffset", "<u4"), ("level", "<u2"), ("fmt", "<u2"), ("pad", "<u4")]))   # stride 12
    tri_recs[1] = (1, 1, abi.OMM_FORMAT_2_STATE, 0)
    tri_recs[0] = (0, 0, abi.OMM_FORMAT_2_STATE, 0)
    data = np.array([0b1, 0b0010], np.uint8)
    mm_idx = np.array([0xFFFF, 0], np.uint16)           # triangle 0: special index -1; triangle 1: record 0 + base 1
    chunks, views = [], []

    def add(b, **kw):
        off = sum(len(c) for c in chunks)
        chunks.append(b + b"\0" * ((-len(b)) % 4))
        views.append(dict(buffer=0, byteOffset=off, byteLength=len(b), **kw))
        return len(views) - 1
    v_pos, v_idx, v_data, v_tri, v_mmi = add(pos.tobytes()), add(idx.tobytes()), add(data.tobytes()), add(tri_recs.tobytes(), byteStride=12), add(mm_idx.tobytes())
    blob = b"".join(chunks)
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
        "extensionsUsed": ["EXT_mesh_opacity_micromap"],
        "extensions": {"EXT_mesh_opacity_micromap": {"micromaps": [
            {"data": v_data, "triangles": v_tri, "usageCounts": [1, 1], "usageLevels": [0, 1], "usageFormats": [1, 1]},
            {"data": v_data, "triangles": v_tri, "usageCounts": [2]}]}},           # second entry: missing fields -> skipped
        "materials": [{"alphaMode": "MASK", "alphaCutoff": 0.5, "doubleSided": True}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": 0,
                                    "extensions": {"EXT_mesh_opacity_micromap": {"micromap": 0, "micromapBaseTriangle": 1, "micromapIndices": 2}}}]}],
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "bufferViews": views,
        "accessors": [{"bufferView": v_pos, "componentType": 5126, "count": 4, "type": "VEC3", "min": [-1, -1, 0], "max": [1, 1, 0]},
                      {"bufferView": v_idx, "componentType": 5123, "count": 6, "type": "SCALAR"},
                      {"bufferView": v_mmi, "componentType": 5123, "count": 2, "type": "SCALAR"}],
    }
    path = tmp_path / "quad_omm.gltf"
    path.write_text(json.dumps(doc))
    scn = scn_mod.load_gltf(str(path))
    assert len(scn.micromaps) == 1 and len(scn.prim_omms) == 1
    assert scn.micromaps[0]["triangles"].tolist() == [(0, 0, 1), (1, 1, 1)] and scn.micromaps[0]["data"].tolist() == [1, 2]
    po = scn.prim_omms[0]
    assert po["baseTriangle"] == 1 and po["indices"].tolist() == [-1, 0] and po["micromap"] == 0
    o = oracle_mod.Oracle()
    o.set_scene(scn)
    rng = np.random.default_rng(4)
    n = 4000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:2] = rng.uniform(-1, 1, (n, 2))
    rays[:, 2] = 2.0
    rays[:, 4:7] = [0, 0, -1]
    rays[:, 7] = 1e30
    seeds = np.arange(n, dtype=np.uint32)
    s = seeds.copy()
    h = o.trace_closest(rays, s)
    hit = h[:, 0] < 1e30
    x, y = rays[:, 0], rays[:, 1]
    # triangle 1 = (v0, v2, v3): u = weight of v2, v = weight of v3; x = -1 + 2u, y = -1 + 2u + 2v; centre micro-triangle: u,v <= 1/2 <= u+v
    u, v = (x + 1) / 2, (y - x) / 2
    expect = (y > x) & (u <= 0.5) & (v <= 0.5) & (u + v >= 0.5)
    margin = np.minimum.reduce([np.abs(y - x), np.abs(u - 0.5), np.abs(v - 0.5), np.abs(u + v - 0.5)]) > 1e-4
    assert np.array_equal(hit[margin], expect[margin]) and np.array_equal(s, seeds) and 0.05 < hit.mean() < 0.3
