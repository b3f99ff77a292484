"""Host side of the animation feed (include/b200pt.h: b200pt_set_animation / b200pt_animate).

Mirrors what SceneAnimationVk prepares for the morph / skinning compute shaders
(src/gltf_scene_animation_vk.cpp:120-260 static task data, :396-494 per-frame data): the task records and the per-frame
joint / normal matrices.  The vertex work itself runs on the device (csrc/animate.cuh); nothing here touches vertices.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class MorphTask:
    """one morphed render primitive (MorphResult + MorphGpuData, gltf_scene_animation_vk.cpp:262-330): base arrays and
    per-target deltas [numTargets, vertexCount, 3]"""
    render_prim: int
    base_positions: np.ndarray
    position_deltas: np.ndarray
    base_normals: Optional[np.ndarray] = None
    base_tangents: Optional[np.ndarray] = None
    normal_deltas: Optional[np.ndarray] = None
    tangent_deltas: Optional[np.ndarray] = None
    mesh: int = -1       # glTF mesh whose `weights` drive the targets (cmdUpdateAnimation :440-458)


@dataclass
class SkinTask:
    """one skinned render primitive (SkinTask + SkinGpuData, gltf_scene_animation_vk.cpp:170-226): base arrays, WEIGHTS_0 [V,4],
    JOINTS_0 [V,4] (int32), the skin's joint count"""
    render_prim: int
    base_positions: np.ndarray
    weights: np.ndarray
    joints: np.ndarray
    num_joints: int
    base_normals: Optional[np.ndarray] = None
    base_tangents: Optional[np.ndarray] = None
    skin: int = -1       # glTF skin index (SkinTask::skinID) and the mesh node the joints are expressed against (refNodeID)
    ref_node: int = -1


def joint_matrices(node_world, joint_nodes, inverse_bind, ref_node):
    """Per-frame skin matrices as cmdUpdateAnimation computes them (gltf_scene_animation_vk.cpp:470-484):
        jointMat[i]  = inverse(world[refNode]) * world[joint_i] * inverseBind[i]   (identity when the skin has fewer)
        normalMat[i] = transpose(inverse(mat3(jointMat[i])))
    node_world / inverse_bind: mathematical 4x4 matrices (row, column).  Returns (jointMats [J,4,4], normalMats [J,3,3]) in the
    glm byte order the C-ABI takes: element [j, c, r] = column c, row r.  (The reference evaluates these few products with glm in
    fp32; here they are evaluated in fp64 and rounded once -- the kernels take whatever matrices the host hands them.)"""
    inv_ref = np.linalg.inv(np.asarray(node_world[ref_node], np.float64))
    jm, nm = [], []
    for i, jn in enumerate(joint_nodes):
        ib = np.asarray(inverse_bind[i], np.float64) if i < len(inverse_bind) else np.eye(4)
        m = inv_ref @ np.asarray(node_world[jn], np.float64) @ ib
        jm.append(m.T)                               # column-major bytes
        nm.append(np.linalg.inv(m[:3, :3]))         # transpose(inverse(m3)) stored column-major: element [c, r] = inverse[c, r]
    return np.asarray(jm, np.float32), np.asarray(nm, np.float32)


def topo_levels(parents):
    """BFS levels of a node forest: (topoNodeOrder int32[numNodes], levelOffsets uint32[numLevels + 1]) -- what the reference
    uploads for world_matrix_propagate.comp.slang (one dispatch per level, parents always in an earlier level)."""
    parents = np.asarray(parents, np.int64)
    depth = np.full(len(parents), -1, np.int64)
    for n in range(len(parents)):
        chain = []
        k = n
        while depth[k] < 0 and parents[k] >= 0:
            chain.append(k)
            k = parents[k]
            if len(chain) > len(parents):
                raise ValueError("node hierarchy has a cycle")
        if depth[k] < 0:
            depth[k] = 0
        d = depth[k]
        for c in reversed(chain):
            d += 1
            depth[c] = d
    order = np.argsort(depth, kind="stable").astype(np.int32)
    counts = np.bincount(depth, minlength=int(depth.max()) + 1)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    return order, offsets


# ---- what the loader's scene graph yields (vk_gltf_renderer_b200.scene.load_gltf fills Scene.graph / morph_prims / skin_prims) ----------
def tasks_from_scene(scn):
    """AnimationSystem::parseMorphTargets / parseSkinTasks (src/gltf_scene_animation.cpp:150-262, 270-316): one MorphTask per render
    primitive with morph targets (base normals / tangents only when a target moves them), one SkinTask per UNIQUE skinned render
    primitive (first render node wins), with the static base arrays taken from the primitive as loaded."""
    morphs, skins = [], []
    for pid, mp in sorted(scn.morph_prims.items()):
        prim = scn.render_prims[pid]
        morphs.append(MorphTask(pid, prim["positions"].copy(), mp["position_deltas"], mesh=mp["mesh"],
                                base_normals=prim["normals"].copy() if (mp["normal_deltas"] is not None and prim["normals"] is not None) else None,
                                base_tangents=prim["tangents"].copy() if (mp["tangent_deltas"] is not None and prim["tangents"] is not None) else None,
                                normal_deltas=mp["normal_deltas"], tangent_deltas=mp["tangent_deltas"]))
    seen = set()
    for rn, (node_id, skin_id, _) in zip(scn.render_nodes, scn.graph["render_nodes"] if scn.graph else []):
        pid = rn["renderPrimID"]
        if skin_id < 0 or pid in seen or pid not in scn.skin_prims or skin_id >= len(scn.graph["skins"]):
            continue
        seen.add(pid)
        prim, sp = scn.render_prims[pid], scn.skin_prims[pid]
        skins.append(SkinTask(pid, prim["positions"].copy(), sp["weights"], sp["joints"], len(scn.graph["skins"][skin_id]["joints"]),
                              base_normals=None if prim["normals"] is None else prim["normals"].copy(),
                              base_tangents=None if prim["tangents"] is None else prim["tangents"].copy(), skin=skin_id, ref_node=node_id))
    return morphs, skins


def node_hierarchy(scn):
    """(parents, RenderNodeGpuMapping triples, instance matrices [R,4,4] glm-ordered) for PathTracer.set_node_hierarchy"""
    g = scn.graph
    mappings = [(node_id, rn["materialID"], rn["renderPrimID"]) for rn, (node_id, _, _) in zip(scn.render_nodes, g["render_nodes"])]
    inst = np.ascontiguousarray(np.asarray([lm for _, _, lm in g["render_nodes"]], np.float64).reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float32))
    return g["parents"], mappings, inst


def world_matrices(parents, locals_):
    """fp64 world matrices of the node graph (the host's view; the device propagates its own in fp32)"""
    locals_ = np.asarray(locals_, np.float64)
    world = [None] * len(parents)

    def get(n):
        if world[n] is None:
            world[n] = locals_[n] if parents[n] < 0 else get(int(parents[n])) @ locals_[n]
        return world[n]
    return np.asarray([get(n) for n in range(len(parents))])


def frame_inputs(scn, morph_tasks, skin_tasks, locals_=None, mesh_weights=None):
    """The per-frame data of one cmdUpdateAnimation (src/gltf_scene_animation_vk.cpp:430-494) for the given node-local matrices
    (default: the asset's rest pose) and mesh weights (default: mesh.weights): morph weights per morph task, joint / normal matrices per
    skin task.  (Sampling animation channels into those locals / weights is the reference's AnimationSystem and stays out of scope.)"""
    g = scn.graph
    world = world_matrices(g["parents"], g["locals"] if locals_ is None else locals_)
    weights = []
    for t in morph_tasks:
        w = (mesh_weights or {}).get(t.mesh, g["mesh_weights"].get(t.mesh, np.zeros(0, np.float32)))
        full = np.zeros(t.position_deltas.shape[0], np.float32)          # fewer weights than targets: the rest stay 0 (:446-452)
        full[:min(len(w), len(full))] = np.asarray(w, np.float32)[:len(full)]
        weights.append(full)
    jms, nms = [], []
    for t in skin_tasks:
        sk = g["skins"][t.skin]
        ibm = [] if sk["ibm"] is None else list(sk["ibm"])
        jm, nm = joint_matrices(world, sk["joints"], ibm, t.ref_node)
        jms.append(jm)
        nms.append(nm)
    return weights, jms, nms


# ---- the caller of the feed: sampling an animation into node-local matrices and mesh weights ---------------------------------------
def _slerp(q1, q2, t):
    """glm::slerp(x, y, a) on (x, y, z, w) quaternions: shortest path, linear blend when the two are (nearly) parallel"""
    q1, q2 = np.asarray(q1, np.float64), np.asarray(q2, np.float64)
    c = float(np.dot(q1, q2))
    if c < 0.0:
        q2, c = -q2, -c
    if c > 1.0 - 1.1920929e-07:
        return q1 + (q2 - q1) * t
    ang = np.arccos(c)
    return (np.sin((1.0 - t) * ang) * q1 + np.sin(t * ang) * q2) / np.sin(ang)


def _cubic(values, t, key_delta, index):
    """computeCubicInterpolation (src/gltf_scene_animation.cpp:498-517): glTF Hermite spline over [in-tangent, value, out-tangent] triplets"""
    t2, t3 = t * t, t * t * t
    c_v1 = -2 * t3 + 3 * t2
    c_v0 = 1 - c_v1
    c_a = key_delta * (t3 - t2)
    c_b = key_delta * (t3 - 2 * t2 + t)
    prev, nxt = index * 3, (index + 1) * 3
    return values[prev + 1] * c_v0 + values[nxt + 0] * c_a + values[prev + 2] * c_b + values[nxt + 1] * c_v1


def sample_animation(scn, animation=0, time=0.0):
    """AnimationSystem::updateAnimation / processAnimationChannel (src/gltf_scene_animation.cpp:352-479) evaluated from the asset's rest
    pose: every translation / rotation / scale / weights channel whose keyframe range contains `time` overwrites its node's value --
    LINEAR (slerp for rotations, :530-583), STEP (:598-628; weights are left alone there) and CUBICSPLINE (:640-688; rotation
    re-normalised) -- then the node-local matrices are rebuilt (T * R * S; a node given as a matrix and not animated keeps it).
    KHR_animation_pointer channels belong to the material / light side of the reference and are not part of this feed.
    Returns (locals f64 [N, 4, 4], mesh_weights {mesh: f32[T]}) for frame_inputs / PathTracer.update_node_matrices."""
    from .scene import _trs
    g = scn.graph
    trs = [dict(t) for t in g["trs"]]
    weights = {m: w.copy() for m, w in g["mesh_weights"].items()}
    animated = set()
    if 0 <= animation < len(g["animations"]):
        an = g["animations"][animation]
        for ch in an["channels"]:
            node, path = ch["node"], ch["path"]
            if not (0 <= node < len(trs)) or path not in ("translation", "rotation", "scale", "weights"):
                continue
            sm = an["samplers"][ch["sampler"]]
            inp = sm["inputs"]
            if len(inp) < 2:
                continue
            k = int(np.searchsorted(inp, np.float32(time), side="right"))     # first key strictly after `time`
            if k == 0:
                continue
            i = min(k - 1, len(inp) - 2)
            t0, t1 = float(inp[i]), float(inp[i + 1])
            if time < t0 or time > t1:
                continue
            dt = t1 - t0
            f = 0.0 if abs(dt) < 1.1920929e-07 else min(max((time - t0) / dt, 0.0), 1.0)
            out = sm["outputs"]
            mode = sm["interpolation"]
            if path == "weights":
                mesh = trs[node]["mesh"]
                if mesh < 0 or mode != "LINEAR":
                    continue
                nt = out.size // len(inp)
                o = out.reshape(len(inp), nt)
                weights[mesh] = ((1.0 - f) * o[i] + f * o[i + 1]).astype(np.float32)
                continue
            o = out.reshape(-1, 4 if path == "rotation" else 3).astype(np.float64)
            if mode == "STEP":
                v = o[i]
            elif mode == "CUBICSPLINE":
                if len(o) <= (i + 1) * 3 + 1:
                    continue
                v = _cubic(o, f, dt, i)
                if path == "rotation":
                    v = v / np.linalg.norm(v)
            else:
                if path == "rotation":
                    v = _slerp(o[i], o[i + 1], f)
                    v = v / np.linalg.norm(v)
                else:
                    v = (1.0 - f) * o[i] + f * o[i + 1]
            trs[node][path] = v.tolist()
            animated.add(node)
    locals_ = g["locals"].copy()
    for n in animated:
        t = trs[n]
        locals_[n] = _trs(np.asarray(t["translation"], np.float64), np.asarray(t["rotation"], np.float64), np.asarray(t["scale"], np.float64))
    return locals_, weights
