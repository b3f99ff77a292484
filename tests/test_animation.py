"""Animation feed without a GPU: the numpy restatement of the reference's morph / skinning compute shaders
(oracle/animation.py; shaders/morph.comp.slang:29-70, shaders/skinning.comp.slang:27-70) against analytic known answers, and
the DEVICE source (csrc/animate.cuh, the bodies of k_morph / k_skin) compiled for the host and compared with it bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vk_gltf_renderer_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
sys.path.insert(0, ROOT)


def _glm(m):
    """mathematical 4x4 / 3x3 -> glm byte order [c, r]"""
    return np.asarray(m, np.float64).T.astype(np.float32)


def test_skinning_known_answers():
    """one joint with weight 1 is a rigid transform of positions, the inverse-transpose on normals, the upper 3x3 on tangents
    (w kept); an identity skeleton reproduces the base mesh; influences with weight 0, a negative joint or a joint beyond the
    skin are skipped (skinning.comp.slang:49-52)."""
    from oracle import animation as A
    rng = np.random.default_rng(0)
    V = 200
    p = rng.normal(size=(V, 3)).astype(np.float32)
    n = rng.normal(size=(V, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    t = np.concatenate([np.cross(n, [0.3, 0.5, 0.8]), np.where(rng.random((V, 1)) < 0.5, -1.0, 1.0)], 1).astype(np.float32)
    c, s = np.cos(0.9), np.sin(0.9)
    M = np.array([[c, -s, 0, 0.5], [s, c, 0, -0.25], [0, 0, 1.7, 0.1], [0, 0, 0, 1.0]])
    jm = np.stack([_glm(M), _glm(np.eye(4))])
    nm = np.stack([_glm(np.linalg.inv(M[:3, :3]).T), _glm(np.eye(3))])
    w = np.zeros((V, 4), np.float32)
    j = np.zeros((V, 4), np.int32)
    w[:, 0] = 1.0
    w[:, 1], j[:, 1] = 0.0, 1       # weight 0: skipped
    w[:, 2], j[:, 2] = 0.7, -1      # negative joint: skipped
    w[:, 3], j[:, 3] = 0.4, 2       # beyond the skin (2 joints): skipped
    sp, sn, st = A.skin(p, n, t, w, j, jm, nm)
    ref_p = (p.astype(np.float64) @ M[:3, :3].T + M[:3, 3])
    assert np.allclose(sp, ref_p, rtol=0, atol=2e-6)
    ref_n = n.astype(np.float64) @ np.linalg.inv(M[:3, :3])
    ref_n /= np.linalg.norm(ref_n, axis=1, keepdims=True)
    assert np.allclose(sn, ref_n, atol=2e-6)
    ref_t = t[:, :3].astype(np.float64) @ M[:3, :3].T
    ref_t /= np.linalg.norm(ref_t, axis=1, keepdims=True)
    assert np.allclose(st[:, :3], ref_t, atol=2e-6) and np.array_equal(st[:, 3], t[:, 3])
    # identity skeleton, weights summing to 1 over two joints: the base mesh up to rounding of the weighted sum
    w2 = np.zeros((V, 4), np.float32)
    w2[:, 0], w2[:, 1] = 0.25, 0.75
    j2 = np.zeros((V, 4), np.int32)
    j2[:, 1] = 1
    ident = np.stack([_glm(np.eye(4))] * 2), np.stack([_glm(np.eye(3))] * 2)
    sp2, sn2, _ = A.skin(p, n, t, w2, j2, *ident)
    assert np.allclose(sp2, p, atol=1e-6) and np.allclose(sn2, n, atol=1e-6)


def test_morph_known_answers():
    """weights of 0 skip their target (morph.comp.slang:45-47): all-zero weights return the base positions bit for bit and
    normalised base normals; one target with weight 1 adds exactly its deltas; tangent.w passes through."""
    from oracle import animation as A
    rng = np.random.default_rng(1)
    V = 150
    p = rng.normal(size=(V, 3)).astype(np.float32)
    n = (rng.normal(size=(V, 3)) * 3).astype(np.float32)
    t = rng.normal(size=(V, 4)).astype(np.float32)
    dp, dn, dt = (rng.normal(size=(3, V, 3)).astype(np.float32) for _ in range(3))
    pos, nrm, tan = A.morph(p, n, t, dp, dn, dt, [0.0, 0.0, 0.0])
    assert np.array_equal(pos, p)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-6) and np.allclose(nrm, n / np.linalg.norm(n, axis=1, keepdims=True), atol=1e-6)
    assert np.array_equal(tan[:, 3], t[:, 3])
    pos, nrm, tan = A.morph(p, n, t, dp, dn, dt, [0.0, 1.0, 0.0])
    assert np.array_equal(pos, p + dp[1])
    pos, nrm, tan = A.morph(p, None, None, dp, dn, dt, [0.5, 0.0, -2.0])
    assert nrm is None and tan is None
    assert np.array_equal(pos, (p + np.float32(0.5) * dp[0]) + np.float32(-2.0) * dp[2])


def _write_task(path, kind, V, K, bp, bn, bt, rest, dn=None, dt=None):
    with open(path, "wb") as f:
        f.write(np.array([kind, V, K, bn is not None, bt is not None, dn is not None, dt is not None], np.uint32).tobytes())
        for a in (bp, bn, bt):
            if a is not None:
                f.write(np.ascontiguousarray(a, np.float32).tobytes())
        for a in rest:
            if a is not None:
                f.write(np.ascontiguousarray(a).tobytes())


def test_device_animation_source_matches_the_oracle(tmp_path):
    """csrc/animate.cuh compiled for the host (tools/host_animate_check.cpp): the morph and skin bodies the sm_100a kernels run
    give the same bits as oracle/animation.py on every task and pose of the animated test scene, including the morph -> skin
    composition of the banner."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    from oracle import animation as A
    from vk_gltf_renderer_b200 import synth
    exe = str(tmp_path / "host_animate_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(CSRC, "tools", "host_animate_check.cpp")])
    scn, morphs, skins, pose = synth.synth_animated()

    def run(kind, V, K, bp, bn, bt, rest, dn=None, dt=None):
        _write_task(str(tmp_path / "task.bin"), kind, V, K, bp, bn, bt, rest, dn, dt)
        subprocess.check_call([exe, str(tmp_path / "task.bin"), str(tmp_path / "out.bin")], stdout=subprocess.DEVNULL)
        raw = np.fromfile(str(tmp_path / "out.bin"), np.float32)
        pos, o = raw[:V * 3].reshape(V, 3), V * 3
        nrm = tan = None
        if bn is not None:
            nrm, o = raw[o:o + V * 3].reshape(V, 3), o + V * 3
        if bt is not None:
            tan = raw[o:o + V * 4].reshape(V, 4)
        return pos, nrm, tan

    def same(a, b):
        return (a is None and b is None) or np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
    for k in range(3):
        mw, jm, nm = pose(k)
        morphed = {}
        for t, w in zip(morphs, mw):
            V = len(t.base_positions)
            got = run(0, V, len(w), t.base_positions, t.base_normals, t.base_tangents, [t.position_deltas, t.normal_deltas, t.tangent_deltas, w],
                      t.normal_deltas, t.tangent_deltas)
            ref = A.morph(t.base_positions, t.base_normals, t.base_tangents, t.position_deltas, t.normal_deltas, t.tangent_deltas, w)
            assert all(same(g, r) for g, r in zip(got, ref)), ("morph", k, t.render_prim)
            morphed[t.render_prim] = ref
        for t, m, n_ in zip(skins, jm, nm):
            bp, bn, bt = morphed.get(t.render_prim, (t.base_positions, t.base_normals, t.base_tangents))
            V = len(bp)
            got = run(1, V, t.num_joints, bp, bn, bt, [np.asarray(t.weights, np.float32), np.asarray(t.joints, np.int32), m, n_])
            ref = A.skin(bp, bn, bt, t.weights, t.joints, m, n_)
            assert all(same(g, r) for g, r in zip(got, ref)), ("skin", k, t.render_prim)
            assert np.isfinite(ref[0]).all()


def _glm_stack(ms):
    return np.ascontiguousarray(np.asarray(ms, np.float64).transpose(0, 2, 1).astype(np.float32))


def test_rigid_feed_known_answers_and_device_source(tmp_path):
    """World-matrix propagation and render-node update (oracle/animation.py propagate / render_nodes / mat4_inverse;
    shaders/world_matrix_propagate.comp.slang:27-42, update_render_instances.comp.slang:42-66): BFS levels put every parent before
    its children, the propagated matrices equal the fp64 chain products, inverse * matrix = identity; and the DEVICE source
    (csrc/animate.cuh propagateNode / updateRenderNode) compiled for the host gives the same bits."""
    from oracle import animation as A
    from vk_gltf_renderer_b200 import synth
    from vk_gltf_renderer_b200.animation import topo_levels
    scn, parents, mappings, inst, pose = synth.synth_hierarchy()
    order, offsets = topo_levels(parents)
    level = {int(n): l for l in range(len(offsets) - 1) for n in order[offsets[l]:offsets[l + 1]]}
    assert sorted(order.tolist()) == list(range(len(parents))) and all(p < 0 or level[int(p)] < level[n] for n, p in enumerate(parents))
    assert len(offsets) - 1 == 4
    exe = None
    if os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        exe = str(tmp_path / "host_animate_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), "-o", exe,
                               os.path.join(CSRC, "tools", "host_animate_check.cpp")])
    for k in range(3):
        loc = pose(k)
        world = A.propagate(_glm_stack(loc), parents, order, offsets)
        for n in range(len(parents)):
            m, p = loc[n], parents[n]
            while p >= 0:
                m, p = loc[p] @ m, parents[p]
            assert np.allclose(world[n].T, m, atol=1e-5), n
        o2w, w2o = A.render_nodes(world, mappings, _glm_stack(inst))
        for i, (node, _, _) in enumerate(mappings):
            assert np.allclose(o2w[i].T.astype(np.float64), inst[i] @ world[node].T.astype(np.float64), atol=1e-5)   # the shader's literal order
            assert np.allclose(w2o[i].T.astype(np.float64) @ o2w[i].T.astype(np.float64), np.eye(4), atol=1e-5)
        if exe:
            with open(tmp_path / "task.bin", "wb") as f:
                f.write(np.array([2, len(parents), len(mappings), len(offsets) - 1, 1, 0, 0], np.uint32).tobytes())
                f.write(parents.astype(np.int32).tobytes() + order.astype(np.int32).tobytes() + offsets.astype(np.uint32).tobytes())
                f.write(np.array([[n_, 0, m_, p_] for n_, m_, p_ in mappings], np.int32).tobytes())
                f.write(_glm_stack(loc).tobytes() + _glm_stack(inst).tobytes())
            subprocess.check_call([exe, str(tmp_path / "task.bin"), str(tmp_path / "out.bin")])
            raw = np.fromfile(str(tmp_path / "out.bin"), np.uint8)
            nw = len(parents) * 64
            got_world = raw[:nw].view(np.float32).reshape(-1, 4, 4)
            recs = raw[nw:].reshape(len(mappings), 136)
            assert np.array_equal(got_world.view(np.uint32), world.view(np.uint32))
            assert np.array_equal(recs[:, :64].copy().view(np.uint32).reshape(-1, 4, 4), o2w.view(np.uint32))
            assert np.array_equal(recs[:, 64:128].copy().view(np.uint32).reshape(-1, 4, 4), w2o.view(np.uint32))
            assert np.array_equal(recs[:, 128:].copy().view(np.int32), np.array([[m_, p_] for _, m_, p_ in mappings], np.int32))


def _write_skinned_gltf(path):
    """A hand-written asset: a strip of 8 vertices along +y (0..3) skinned to a chain of two joints (joint B at y = 1.5, child of joint A),
    one morph target (+x bulge) with mesh.weights = [0.25], NORMAL present; joints under an armature node that is translated, the mesh node
    elsewhere (glTF ignores the skinned mesh node's own transform: inverse(meshWorld) * jointWorld * IBM)."""
    import json
    ys = np.repeat(np.linspace(0.0, 3.0, 4), 2)
    pos = np.stack([np.tile([-0.2, 0.2], 4), ys, np.zeros(8)], 1).astype(np.float32)
    nrm = np.tile(np.float32([0, 0, 1]), (8, 1))
    idx = np.uint16([[0, 1, 3], [0, 3, 2], [2, 3, 5], [2, 5, 4], [4, 5, 7], [4, 7, 6]]).reshape(-1)
    wB = np.clip((ys - 1.0) / 1.0, 0, 1).astype(np.float32)
    weights = np.stack([1 - wB, wB, np.zeros(8), np.zeros(8)], 1).astype(np.float32)
    joints = np.tile(np.uint8([0, 1, 0, 0]), (8, 1))
    delta = np.stack([0.5 * np.sin(ys), np.zeros(8), np.zeros(8)], 1).astype(np.float32)
    arm = np.eye(4)
    arm[:3, 3] = [2.0, 0.5, -1.0]
    jb = np.eye(4)
    jb[1, 3] = 1.5
    ibm = np.stack([np.linalg.inv(arm), np.linalg.inv(arm @ jb)]).transpose(0, 2, 1).astype(np.float32)   # glm column-major
    chunks, views, accessors = [], [], []

    def add(arr, ctype, atype, target=None, minmax=False):
        raw = np.ascontiguousarray(arr).tobytes()
        off = sum(len(c) for c in chunks)
        pad = (-len(raw)) % 4
        chunks.append(raw + b"\\0" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(raw)}
        if target:
            v["target"] = target
        views.append(v)
        a = {"bufferView": len(views) - 1, "componentType": ctype, "count": len(arr) if atype != "SCALAR" else int(np.asarray(arr).size), "type": atype}
        if minmax:
            a["min"], a["max"] = np.asarray(arr).min(0).tolist(), np.asarray(arr).max(0).tolist()
        accessors.append(a)
        return len(accessors) - 1
    a_pos = add(pos, 5126, "VEC3", 34962, True)
    a_nrm = add(nrm, 5126, "VEC3", 34962)
    a_idx = add(idx, 5123, "SCALAR", 34963)
    a_w = add(weights, 5126, "VEC4", 34962)
    a_j = add(joints, 5121, "VEC4", 34962)
    a_d = add(delta, 5126, "VEC3", None, True)
    a_ibm = add(ibm.reshape(2, 16), 5126, "MAT4")
    # one animation: jointB rotates 0 -> 90 degrees about z (LINEAR), the mesh weights go 0.25 -> 1 (LINEAR), the armature translates
    # along a Hermite spline (CUBICSPLINE: in-tangent, value, out-tangent per key), jointA's scale steps (STEP)
    h = np.float32(np.sqrt(0.5))
    a_t2 = add(np.float32([0.0, 1.0]), 5126, "SCALAR", None, True)
    a_rot = add(np.float32([[0, 0, 0, 1], [0, 0, h, h]]), 5126, "VEC4")
    a_wgt = add(np.float32([0.25, 1.0]), 5126, "SCALAR")
    a_spl = add(np.float32([[0, 0, 0], [2.0, 0.5, -1.0], [1, 0, 0], [0, 2, 0], [2.0, 1.5, -1.0], [0, 0, 0]]), 5126, "VEC3")
    a_t3 = add(np.float32([0.0, 0.5, 1.0]), 5126, "SCALAR", None, True)
    a_scl = add(np.float32([[1, 1, 1], [1, 2, 1], [1, 3, 1]]), 5126, "VEC3")
    blob = b"".join(chunks)
    with open(path.replace(".gltf", ".bin"), "wb") as f:
        f.write(blob)
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 1]}],
           "nodes": [{"name": "meshNode", "mesh": 0, "skin": 0, "translation": [7.0, 7.0, 7.0]},
                     {"name": "armature", "translation": [2.0, 0.5, -1.0], "children": [2]},
                     {"name": "jointA", "children": [3]},
                     {"name": "jointB", "translation": [0.0, 1.5, 0.0]}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "WEIGHTS_0": a_w, "JOINTS_0": a_j}, "indices": a_idx,
                                       "targets": [{"POSITION": a_d}], "material": 0}], "weights": [0.25]}],
           "skins": [{"joints": [2, 3], "inverseBindMatrices": a_ibm}],
           "animations": [{"name": "bend", "samplers": [{"input": a_t2, "output": a_rot, "interpolation": "LINEAR"}, {"input": a_t2, "output": a_wgt},
                                                         {"input": a_t2, "output": a_spl, "interpolation": "CUBICSPLINE"},
                                                         {"input": a_t3, "output": a_scl, "interpolation": "STEP"}],
                           "channels": [{"sampler": 0, "target": {"node": 3, "path": "rotation"}}, {"sampler": 1, "target": {"node": 0, "path": "weights"}},
                                        {"sampler": 2, "target": {"node": 1, "path": "translation"}}, {"sampler": 3, "target": {"node": 2, "path": "scale"}},
                                        {"sampler": 0, "target": {"node": 3, "path": "pointer"}}]}],
           "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.7, 0.6, 1.0]}, "doubleSided": True}],
           "buffers": [{"uri": os.path.basename(path).replace(".gltf", ".bin"), "byteLength": len(blob)}], "bufferViews": views, "accessors": accessors}
    with open(path, "w") as f:
        json.dump(doc, f)
    return pos, delta, weights, arm, jb


def test_loader_builds_the_animation_tasks_of_a_skinned_morphed_asset(tmp_path):
    """load_gltf gathers what AnimationSystem::parseMorphTargets / parseSkinTasks cache (src/gltf_scene_animation.cpp:150-316) and the node
    graph; animation.tasks_from_scene / frame_inputs turn it into the C-ABI feed.  At the rest pose the joint matrices cancel the inverse
    bind matrices (the skinned mesh node's own translation drops out) and the deformed mesh is base + 0.25 * target; bending joint B by 90
    degrees moves the fully B-weighted vertices around B's origin."""
    from oracle import animation as A
    from vk_gltf_renderer_b200 import animation as anim, scene
    path = str(tmp_path / "skinned.gltf")
    pos, delta, weights, arm, jb = _write_skinned_gltf(path)
    scn = scene.load_gltf(path)
    assert len(scn.render_nodes) == 1 and scn.graph["parents"].tolist() == [-1, -1, 1, 2]
    assert scn.graph["render_nodes"][0][:2] == (0, 0)                       # refNodeID 0, skinID 0
    morphs, skins = anim.tasks_from_scene(scn)
    assert len(morphs) == 1 and len(skins) == 1 and skins[0].num_joints == 2 and skins[0].ref_node == 0 and morphs[0].mesh == 0
    assert np.array_equal(skins[0].weights, weights) and skins[0].joints.dtype == np.int32 and skins[0].joints[:, 1].tolist() == [1] * 8
    assert morphs[0].base_normals is None                                   # no target moves the normals (:227-247)
    parents, mappings, inst = anim.node_hierarchy(scn)
    assert mappings == [(0, 0, 0)] and np.array_equal(inst[0], np.eye(4, dtype=np.float32))
    order, offsets = anim.topo_levels(parents)
    assert offsets.tolist() == [0, 2, 3, 4]
    # rest pose
    mw, jm, nm = anim.frame_inputs(scn, morphs, skins)
    assert np.allclose(mw[0], [0.25]) and np.allclose(jm[0][:, :, :], np.stack([np.linalg.inv(scn.graph["locals"][0]).T] * 2), atol=1e-6)
    ref = scene.load_gltf(path)
    A.apply(ref, morphs, skins, mw, jm, nm)
    mesh_inv = np.linalg.inv(scn.graph["locals"][0])
    expect = (pos + np.float32(0.25) * delta).astype(np.float64) @ mesh_inv[:3, :3].T + mesh_inv[:3, 3]
    assert np.allclose(ref.render_prims[0]["positions"], expect, atol=1e-5)
    # the render node's world matrix puts it back: world position = base + 0.25 * delta, whatever the mesh node's translation is
    o2w = np.asarray(ref.render_nodes[0]["objectToWorld"], np.float64).reshape(4, 4).T
    world_pos = ref.render_prims[0]["positions"].astype(np.float64) @ o2w[:3, :3].T + o2w[:3, 3]
    assert np.allclose(world_pos, pos + 0.25 * delta, atol=1e-5)
    # bend joint B by 90 degrees about z
    locals_ = scn.graph["locals"].copy()
    rz = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    locals_[3] = jb @ rz
    mw, jm, nm = anim.frame_inputs(scn, morphs, skins, locals_, mesh_weights={0: [0.0]})
    bent = scene.load_gltf(path)
    A.apply(bent, morphs, skins, mw, jm, nm)
    world_pos = bent.render_prims[0]["positions"].astype(np.float64) @ o2w[:3, :3].T + o2w[:3, 3]
    top = weights[:, 1] == 1.0                                              # vertices fully on joint B
    # world = jointWorld_new * IBM * p = T(a) T(b) Rz T(-b) T(-a) p: a rotation about joint B's bind-pose origin a + b
    pivot = (arm @ jb)[:3, 3]
    assert np.allclose(world_pos[top], pivot + (pos[top] - pivot) @ rz[:3, :3].T, atol=1e-5)
    assert np.allclose(world_pos[weights[:, 0] == 1.0], pos[weights[:, 0] == 1.0], atol=1e-5)   # joint A's vertices stay
    n = bent.render_prims[0]["normals"]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)


def test_animation_sampler_feeds_the_tasks(tmp_path):
    """animation.sample_animation (AnimationSystem::updateAnimation / processAnimationChannel, src/gltf_scene_animation.cpp:352-688): LINEAR
    slerp / lerp, STEP, CUBICSPLINE with the glTF Hermite basis, weights channels, times outside a sampler's range leave the node alone;
    its locals and mesh weights drive frame_inputs, and the oracle's skinning then puts the fully B-weighted vertices where a 45 degree
    bend about the moved joint puts them."""
    from oracle import animation as A
    from vk_gltf_renderer_b200 import animation as anim, scene
    path = str(tmp_path / "skinned.gltf")
    pos, delta, weights, arm, jb = _write_skinned_gltf(path)
    scn = scene.load_gltf(path)
    assert len(scn.graph["animations"]) == 1 and len(scn.graph["animations"][0]["channels"]) == 5
    rest, w_rest = anim.sample_animation(scn, 0, -1.0)                     # before the first key: nothing is animated
    assert np.array_equal(rest, scn.graph["locals"]) and np.allclose(w_rest[0], [0.25])
    loc, w = anim.sample_animation(scn, 0, 0.5)
    c = np.cos(np.pi / 4)
    assert np.allclose(loc[3][:3, :3], [[c, -c, 0], [c, c, 0], [0, 0, 1]], atol=1e-6) and np.allclose(loc[3][:3, 3], [0, 1.5, 0])   # slerp half way
    assert np.allclose(w[0], [0.625])
    # Hermite: p(t) = (2t^3 - 3t^2 + 1) p0 + dt (t^3 - 2t^2 + t) b0 + (-2t^3 + 3t^2) p1 + dt (t^3 - t^2) a1 at t = 0.5, dt = 1
    p0, b0, a1, p1 = np.float64([2, 0.5, -1]), np.float64([1, 0, 0]), np.float64([0, 2, 0]), np.float64([2, 1.5, -1])
    assert np.allclose(loc[1][:3, 3], 0.5 * p0 + 0.125 * b0 + 0.5 * p1 - 0.125 * a1, atol=1e-6)
    assert np.allclose(np.diag(loc[2])[:3], [1, 2, 1])                      # STEP: the key at 0.5 holds until 1.0
    assert np.allclose(np.diag(anim.sample_animation(scn, 0, 0.49)[0][2])[:3], [1, 1, 1])
    late, _ = anim.sample_animation(scn, 0, 2.0)                            # past the last key: outside every range
    assert np.array_equal(late, scn.graph["locals"])
    # through the feed: joint B bent by 45 degrees about its (moved, stretched) origin
    morphs, skins = anim.tasks_from_scene(scn)
    mw, jm, nm = anim.frame_inputs(scn, morphs, skins, loc, {0: np.zeros(1, np.float32)})
    bent = scene.load_gltf(path)
    A.apply(bent, morphs, skins, mw, jm, nm)
    o2w = np.asarray(bent.render_nodes[0]["objectToWorld"], np.float64).reshape(4, 4).T
    world_pos = bent.render_prims[0]["positions"].astype(np.float64) @ o2w[:3, :3].T + o2w[:3, 3]
    world = anim.world_matrices(scn.graph["parents"], loc)
    top = weights[:, 1] == 1.0
    expect = (np.c_[pos[top], np.ones(top.sum())] @ (world[3] @ np.linalg.inv(arm @ jb)).T)[:, :3]
    assert np.allclose(world_pos[top], expect, atol=1e-5)
